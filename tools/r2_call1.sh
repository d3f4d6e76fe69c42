#!/bin/bash
# Round 2, GPU call 1: baseline of the round + what round 1 wrote after its GPU budget (batched ensembles, flat staging).
# Variant libraries are prebuilt (SPHB_BUILD_TAG) and travel with the snapshot.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_call1
mkdir -p "$OUT"
step() { local name=$1 secs=$2; shift 2
    echo "== $name" | tee -a "$OUT/summary.txt"
    timeout "$secs" "$@" > "$OUT/$name.out" 2> "$OUT/$name.err"
    echo "   exit $?" | tee -a "$OUT/summary.txt"
    tail -n 3 "$OUT/$name.out" | cut -c1-600 | tee -a "$OUT/summary.txt"; }
nvidia-smi > "$OUT/smi.txt"; nproc >> "$OUT/smi.txt"
step 1_tests_default 600 python -m pytest tests -x -q -m gpu
step 1_bench_default 300 python bench.py --steps 20 --warmup 3
step 2_ens256_graphs 300 python bench.py --mode ensemble --channels 256 --steps 20 --warmup 3
step 2_ens256_batched 300 python bench.py --mode ensemble --channels 256 --steps 20 --warmup 3 --batched
step 2_ens512_batched 300 python bench.py --mode ensemble --channels 512 --steps 20 --warmup 3 --batched
step 2_ens1024_batched 300 python bench.py --mode ensemble --channels 1024 --steps 20 --warmup 3 --batched
for v in flat flat2 flat4; do
  export SPHB_LIB_VARIANT=$v
  step 3_tests_$v 600 python -m pytest tests -x -q -m gpu
  step 3_bench_$v 300 python bench.py --steps 20 --warmup 3 --e2e-steps 2 --cpu-baseline-steps 0
done
unset SPHB_LIB_VARIANT
cat "$OUT/summary.txt"
