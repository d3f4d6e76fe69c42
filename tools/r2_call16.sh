#!/bin/bash
# Round 2, GPU call 16 (1 GPU): continuity's own operands requested before its row loop; tiles per CTA of the batched
# ensemble kernels (SPHB_BATCH_TILES_PER_CTA); driver with the version-2 restart header.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_call16
mkdir -p "$OUT"
step() { local name=$1 secs=$2; shift 2
    echo "== $name" | tee -a "$OUT/summary.txt"
    timeout "$secs" "$@" > "$OUT/$name.out" 2> "$OUT/$name.err"
    echo "   exit $?" | tee -a "$OUT/summary.txt"
    tail -n 3 "$OUT/$name.out" | cut -c1-300 | tee -a "$OUT/summary.txt"; }
step 1_tests 900 python -m pytest tests -q -m gpu
step 1_bench 300 python bench.py --steps 50 --warmup 5 --e2e-steps 2 --cpu-baseline-steps 0
for n in 1 2 3 6 12; do
    SPHB_BATCH_TILES_PER_CTA=$n step 2_ens512_t$n 200 python bench.py --mode ensemble --channels 512 --steps 20 --warmup 3
done
python - <<'PY' | tee -a gpurun_out/r2_call16/summary.txt
import json, glob
for f in sorted(glob.glob('gpurun_out/r2_call16/*.out')):
    for l in open(f):
        if l.startswith('{'):
            d = json.loads(l)
            print(f.split('/')[-1], round(d['ms_per_step'], 4), f"{d['value']:.4e}", {k: round(v['ms_per_call'] * 1e3, 1) for k, v in d.get('kernels', {}).items()})
PY
