#!/bin/bash
# Round 2, GPU call 43 (2 GPUs): the whole GPU suite on a 2-GPU box (NCCL / mailbox slab tests included), the default
# bench at N = 1 and, sharded over two slabs, at N = 2; the C host of the sharded channel.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_call43
mkdir -p "$OUT"
step() { local name=$1 secs=$2; shift 2
    echo "== $name" | tee -a "$OUT/summary.txt"
    timeout "$secs" "$@" > "$OUT/$name.out" 2> "$OUT/$name.err"
    echo "   exit $?" | tee -a "$OUT/summary.txt"
    tail -n 3 "$OUT/$name.out" | cut -c1-300 | tee -a "$OUT/summary.txt"; }
export SPHB_ERRLOG=$PWD/$OUT/observed_errors.jsonl
step t_all 900 python -m pytest tests -q -m gpu
unset SPHB_ERRLOG
step b_n1 300 python bench.py --cpu-baseline-steps 0
step b_n2 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2
cat "$OUT/summary.txt"
