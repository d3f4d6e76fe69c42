#!/bin/bash
# Round 2, GPU call 5: list build folded (grid set-up in the continuity epilogue, canonical order + tile descriptors
# inside the walk), wall shear on candidate rows only: GPU suite, bench, launch list.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_call6
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


cat "$OUT/summary.txt"
