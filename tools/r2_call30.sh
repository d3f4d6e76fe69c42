#!/bin/bash
# Round 2, GPU call 30 (1 GPU): suite + default bench on HEAD, ncu launch list and full capture of the step kernels (r2_v3).
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_call30
mkdir -p "$OUT"
step() { local name=$1 secs=$2; shift 2
    echo "== $name" | tee -a "$OUT/summary.txt"
    timeout "$secs" "$@" > "$OUT/$name.out" 2> "$OUT/$name.err"
    echo "   exit $?" | tee -a "$OUT/summary.txt"
    tail -n 3 "$OUT/$name.out" | cut -c1-300 | tee -a "$OUT/summary.txt"; }
export SPHB_ERRLOG=$PWD/$OUT/observed_errors.jsonl
step 1_tests 900 python -m pytest tests -q -m gpu
unset SPHB_ERRLOG
step 1_bench 300 python bench.py
step 3_ncu_launches 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches.csv" python bench.py --steps 3 --warmup 3 --e2e-steps 1 --cpu-baseline-steps 0 --profile-steps 1
step 3_ncu_full 900 ncu --set full --clock-control none --import-source on -k regex:'k_(force_fused|continuity|kgc|density|walk_tile|wall_shear|bin_count|scatter|publish_dt)$|scan_' -s 40 -c 12 -o "$OUT/prof_r2_v3" python bench.py --steps 3 --warmup 3 --e2e-steps 1 --cpu-baseline-steps 0 --profile-steps 1
cat "$OUT/summary.txt"
