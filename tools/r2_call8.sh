#!/bin/bash
# Round 2, GPU call 7: suite with the poison / recovery / dual-auto changes, bench (+ device-side scheduler at low
# viscosity), fixed inner = 10 (BASELINE configs[2] "fixed inner acoustic substeps"), ensemble grids.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_call8
mkdir -p "$OUT"
step() { local name=$1 secs=$2; shift 2
    echo "== $name" | tee -a "$OUT/summary.txt"
    timeout "$secs" "$@" > "$OUT/$name.out" 2> "$OUT/$name.err"
    echo "   exit $?" | tee -a "$OUT/summary.txt"
    tail -n 3 "$OUT/$name.out" | cut -c1-300 | tee -a "$OUT/summary.txt"; }
export SPHB_ERRLOG=$PWD/$OUT/observed_errors.jsonl
step 1_tests 900 python -m pytest tests -q -m gpu
unset SPHB_ERRLOG
step 1_bench 300 python bench.py --steps 20 --warmup 3
step 2_bench_auto 300 python bench.py --steps 20 --warmup 3 --mu 0.005 --dual-auto 16 --e2e-steps 2 --cpu-baseline-steps 0
step 3_ens512 300 python bench.py --mode ensemble --channels 512 --steps 20 --warmup 3
cat "$OUT/summary.txt"
