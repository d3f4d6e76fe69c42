#!/bin/bash
# Round 2, GPU call 36 (1 GPU): A/B of prebuilt library variants (SPHB_LIB_VARIANT) on the default bench --
#   base  = HEAD (CTA-wide tile pipeline)            (default build) = warp-granular tile pipeline (SPHB_WARP_PIPELINE=1)
#   cubic = base + one third-order Newton step       wpc  = warp pipeline + cubic       wpc5 = wpc, density / KGC at 5 CTAs per SM
#   w4 / w5 / w8 = base with the walk compiled for 4 / 5 / 8 resident CTAs
# and the GPU suite on wpc.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_call36
mkdir -p "$OUT"
step() { local name=$1 secs=$2; shift 2
    echo "== $name" | tee -a "$OUT/summary.txt"
    timeout "$secs" "$@" > "$OUT/$name.out" 2> "$OUT/$name.err"
    echo "   exit $?" | tee -a "$OUT/summary.txt"
    tail -n 3 "$OUT/$name.out" | cut -c1-300 | tee -a "$OUT/summary.txt"; }
B="python bench.py --steps 100 --warmup 5 --e2e-steps 2 --cpu-baseline-steps 0"
SPHB_LIB_VARIANT=base step b_base 200 $B
step b_wp 200 $B
SPHB_LIB_VARIANT=wpc step b_wpc 200 $B
SPHB_LIB_VARIANT=cubic step b_cubic 200 $B
SPHB_LIB_VARIANT=wpc5 step b_wpc5 200 $B
for v in w4 w5 w8; do SPHB_LIB_VARIANT=$v step b_$v 200 $B; done
SPHB_LIB_VARIANT=base step b_base2 200 $B
export SPHB_ERRLOG=$PWD/$OUT/observed_errors.jsonl
SPHB_LIB_VARIANT=wpc step t_wpc 900 python -m pytest tests -q -m gpu -x
unset SPHB_ERRLOG
python - <<'PY' | tee -a gpurun_out/r2_call36/summary.txt
import json, glob
for f in sorted(glob.glob('gpurun_out/r2_call36/b_*.out')):
    for l in open(f):
        if l.startswith('{'):
            d = json.loads(l)
            print(f.split('/')[-1], round(d['ms_per_step'], 4), f"{d['value']:.4e}", {k: round(v['ms_per_call'] * 1e3, 1) for k, v in d.get('kernels', {}).items()})
PY
