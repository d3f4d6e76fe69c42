#!/bin/bash
# Round 2, GPU call 3: the persistent, double-buffered bulk-copy pipeline of the gather kernels (ops_csr.cu run_tiles):
# full GPU suite + observed errors, bench of the default build and of the tile-row / occupancy variants.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_call3
mkdir -p "$OUT"
step() { local name=$1 secs=$2; shift 2
    echo "== $name" | tee -a "$OUT/summary.txt"
    timeout "$secs" "$@" > "$OUT/$name.out" 2> "$OUT/$name.err"
    echo "   exit $?" | tee -a "$OUT/summary.txt"
    tail -n 3 "$OUT/$name.out" | cut -c1-400 | tee -a "$OUT/summary.txt"; }
export SPHB_ERRLOG=$PWD/$OUT/observed_errors.jsonl
step 1_tests 900 python -m pytest tests -q -m gpu
unset SPHB_ERRLOG
step 1_bench 300 python bench.py --steps 20 --warmup 3
for v in tr3 rx tr3rx; do
  export SPHB_LIB_VARIANT=$v
  step 2_bench_$v 300 python bench.py --steps 20 --warmup 3 --e2e-steps 2 --cpu-baseline-steps 0
done
export SPHB_LIB_VARIANT=tr3
step 3_tests_tr3 900 python -m pytest tests -q -m gpu -x
unset SPHB_LIB_VARIANT
step 4_ens512 300 python bench.py --mode ensemble --channels 512 --steps 20 --warmup 3
cat "$OUT/summary.txt"
