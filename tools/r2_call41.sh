#!/bin/bash
# Round 2, GPU call 41 (1 GPU): two row words in flight instead of one in the gathers' row loops (pf2 = SPHB_ROW_PF2=1) against HEAD.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_call41
mkdir -p "$OUT"
step() { local name=$1 secs=$2; shift 2
    echo "== $name" | tee -a "$OUT/summary.txt"
    timeout "$secs" "$@" > "$OUT/$name.out" 2> "$OUT/$name.err"
    echo "   exit $?" | tee -a "$OUT/summary.txt"
    tail -n 3 "$OUT/$name.out" | cut -c1-300 | tee -a "$OUT/summary.txt"; }
B="python bench.py --steps 100 --warmup 5 --e2e-steps 2 --cpu-baseline-steps 0"
step b_head 200 $B
SPHB_LIB_VARIANT=pf2 step b_pf2 200 $B
step b_head2 200 $B
SPHB_LIB_VARIANT=pf2 step b_pf2b 200 $B
python - <<'PY' | tee -a gpurun_out/r2_call41/summary.txt
import json, glob
for f in sorted(glob.glob('gpurun_out/r2_call41/[be]_*.out')):
    for l in open(f):
        if l.startswith('{'):
            d = json.loads(l)
            print(f.split('/')[-1], round(d['ms_per_step'], 4), f"{d['value']:.4e}", d.get('gpu_launches'), {k: round(v['ms_per_call'] * 1e3, 1) for k, v in d.get('kernels', {}).items()})
PY
