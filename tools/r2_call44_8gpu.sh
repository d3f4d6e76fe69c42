#!/bin/bash
# Round 2, GPU call 44 (8 GPUs): HEAD lines of the three multi-GPU BASELINE workloads -- weak scaling N = 8 (default),
# configs[4] (DL = 96 sharded over 8 GPUs) and configs[3] (4096-channel ensemble, 512 per GPU).
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_call44
mkdir -p "$OUT"
step() { local name=$1 secs=$2; shift 2
    echo "== $name" | tee -a "$OUT/summary.txt"
    timeout "$secs" "$@" > "$OUT/$name.out" 2> "$OUT/$name.err"
    echo "   exit $?" | tee -a "$OUT/summary.txt"
    tail -n 2 "$OUT/$name.out" | cut -c1-330 | tee -a "$OUT/summary.txt"; }
TR() { echo python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2; }
step 2_scale_n8 200 $(TR 8 29631) bench.py --gpus 8 --steps 60 --warmup 5 --e2e-steps 3
step 3_dl96_n8 200 $(TR 8 29633) bench.py --gpus 8 --DL 12 --steps 30 --warmup 5 --e2e-steps 2
step 4_ens4096_n8 200 $(TR 8 29634) bench.py --gpus 8 --mode ensemble --channels 512 --steps 20 --warmup 3
cat "$OUT/summary.txt"
