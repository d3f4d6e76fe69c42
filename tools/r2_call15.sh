#!/bin/bash
# Round 2, GPU call 15 (1 GPU): suite on HEAD, default bench, BASELINE configs[4] on ONE GPU (DL = 96) and its per-GPU
# share (DL = 12) for the parallel efficiency of call 14, ncu launch list + full capture of the step kernels, launch
# list of one batched ensemble step.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_call15
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
step 2_bench_dl12 400 python bench.py --DL 12 --steps 20 --warmup 3 --e2e-steps 2 --cpu-baseline-steps 0
step 2_bench_dl96 900 python bench.py --DL 96 --steps 10 --warmup 3 --e2e-steps 1 --cpu-baseline-steps 0
step 2_bench_noflush 300 python bench.py --no-flush --steps 30 --warmup 5 --e2e-steps 1 --cpu-baseline-steps 0
step 3_ncu_launches 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches.csv" python bench.py --steps 3 --warmup 3 --e2e-steps 1 --cpu-baseline-steps 0 --profile-steps 1
step 3_ncu_full 900 ncu --set full --clock-control none --import-source on -k regex:'k_(force_fused|continuity|kgc|density|walk_tile|wall_shear|bin_count|scatter|publish_dt)$|scan_' -s 40 -c 12 -o "$OUT/prof_r2_v2" python bench.py --steps 3 --warmup 3 --e2e-steps 1 --cpu-baseline-steps 0 --profile-steps 1
step 4_ens_ncu 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 4450 -c 60 --csv --log-file "$OUT/ens256_launches.csv" python bench.py --mode ensemble --channels 256 --steps 2 --warmup 1
cat "$OUT/summary.txt"
