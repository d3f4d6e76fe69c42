#!/bin/bash
# Round 2, GPU call 4: occupancy / inlining variants of the pipelined gather kernels.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_call4
mkdir -p "$OUT"
step() { local name=$1 secs=$2; shift 2
    echo "== $name" | tee -a "$OUT/summary.txt"
    timeout "$secs" "$@" > "$OUT/$name.out" 2> "$OUT/$name.err"
    echo "   exit $?" | tee -a "$OUT/summary.txt"
    tail -n 3 "$OUT/$name.out" | cut -c1-300 | tee -a "$OUT/summary.txt"; }
step 1_tests 900 python -m pytest tests -q -m gpu -x -k "context or slab"
step 1_bench 300 python bench.py --steps 20 --warmup 3 --e2e-steps 2 --cpu-baseline-steps 0
for v in v2 v3 v4 v5; do
  export SPHB_LIB_VARIANT=$v
  step 2_bench_$v 300 python bench.py --steps 20 --warmup 3 --e2e-steps 2 --cpu-baseline-steps 0
done
unset SPHB_LIB_VARIANT
cat "$OUT/summary.txt"
