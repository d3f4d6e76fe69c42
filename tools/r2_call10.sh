#!/bin/bash
# Round 2, GPU call 10 (1 GPU): suite on HEAD (slab exchange folded: 6 launches), bench, tile-width sweep of the
# single-domain walk (SPHB_TILE_W), launch list of the batched ensemble step.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_call10
mkdir -p "$OUT"
step() { local name=$1 secs=$2; shift 2
    echo "== $name" | tee -a "$OUT/summary.txt"
    timeout "$secs" "$@" > "$OUT/$name.out" 2> "$OUT/$name.err"
    echo "   exit $?" | tee -a "$OUT/summary.txt"
    tail -n 3 "$OUT/$name.out" | cut -c1-300 | tee -a "$OUT/summary.txt"; }
export SPHB_ERRLOG=$PWD/$OUT/observed_errors.jsonl
step 1_tests 900 python -m pytest tests -q -m gpu
unset SPHB_ERRLOG
step 1_bench 300 python bench.py --steps 30 --warmup 5
for w in 5 6 7 8 10 12; do
    SPHB_TILE_W=$w step 2_bench_tw$w 200 python bench.py --steps 20 --warmup 3 --e2e-steps 1 --cpu-baseline-steps 0
done
step 3_ens512 300 python bench.py --mode ensemble --channels 512 --steps 20 --warmup 3
step 4_ens_ncu 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file "$OUT/ens512_launches.csv" \
    python bench.py --mode ensemble --channels 512 --steps 2 --warmup 1
python - <<'PY' | tee -a gpurun_out/r2_call10/summary.txt
import json, glob
for f in sorted(glob.glob('gpurun_out/r2_call10/*_bench*.out')):
    for l in open(f):
        if l.startswith('{'):
            d = json.loads(l)
            print(f.split('/')[-1], round(d['ms_per_step'], 4), {k: round(v['ms_per_call'] * 1e3, 1) for k, v in d.get('kernels', {}).items()})
PY
cat "$OUT/summary.txt"
