#!/bin/bash
# Round 2, GPU call 2: observed errors of every tolerance check, bench with the fp64 roofline, row-loop variants,
# ncu captures (launch list + full set of the step kernels) of the default build.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_call2
mkdir -p "$OUT"
step() { local name=$1 secs=$2; shift 2
    echo "== $name" | tee -a "$OUT/summary.txt"
    timeout "$secs" "$@" > "$OUT/$name.out" 2> "$OUT/$name.err"
    echo "   exit $?" | tee -a "$OUT/summary.txt"
    tail -n 3 "$OUT/$name.out" | cut -c1-400 | tee -a "$OUT/summary.txt"; }
export SPHB_ERRLOG=$PWD/$OUT/observed_errors.jsonl
step 1_tests 900 python -m pytest tests -x -q -m gpu
unset SPHB_ERRLOG
step 1_bench 300 python bench.py --steps 20 --warmup 3
for v in fw fw3 m4 m5; do
  export SPHB_LIB_VARIANT=$v
  step 2_bench_$v 300 python bench.py --steps 20 --warmup 3 --e2e-steps 2 --cpu-baseline-steps 0
done
unset SPHB_LIB_VARIANT
step 3_ncu_launches 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches.csv" python bench.py --steps 3 --warmup 3 --e2e-steps 1 --cpu-baseline-steps 0 --profile-steps 1
step 3_ncu_full 900 ncu --set full --clock-control none --import-source on -k regex:'k_(force_fused|continuity|kgc|density|walk_tile|wall_shear|yrange_grid|bin_count|scatter|cell_canon|tile_setup)' -s 60 -c 14 -o "$OUT/prof_r2_v1" python bench.py --steps 3 --warmup 3 --e2e-steps 1 --cpu-baseline-steps 0 --profile-steps 1
cat "$OUT/summary.txt"
