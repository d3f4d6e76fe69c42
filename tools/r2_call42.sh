#!/bin/bash
# Round 2, GPU call 42 (1 GPU): final capture of the round on HEAD -- default bench line (with the CPU sample), the
# reference arm, ncu launch list and full capture of the step kernels (r2_v4).
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_call42
mkdir -p "$OUT"
step() { local name=$1 secs=$2; shift 2
    echo "== $name" | tee -a "$OUT/summary.txt"
    timeout "$secs" "$@" > "$OUT/$name.out" 2> "$OUT/$name.err"
    echo "   exit $?" | tee -a "$OUT/summary.txt"
    tail -n 3 "$OUT/$name.out" | cut -c1-300 | tee -a "$OUT/summary.txt"; }
step 1_bench 300 python bench.py
step 3_ncu_launches 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches.csv" python bench.py --steps 3 --warmup 3 --e2e-steps 1 --cpu-baseline-steps 0 --profile-steps 1
step 3_ncu_full 900 ncu --set full --clock-control none --import-source on -k regex:'k_(force_fused|continuity|kgc|density|walk_tile|wall_shear|bin_count|scatter)$|scan_' -s 40 -c 11 -o "$OUT/prof_r2_v4" python bench.py --steps 3 --warmup 3 --e2e-steps 1 --cpu-baseline-steps 0 --profile-steps 1
cat "$OUT/summary.txt"
