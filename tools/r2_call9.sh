#!/bin/bash
# Round 2, GPU call 9 (2 GPUs): two-rank slab parity (NCCL and peer-to-peer mailboxes), sharded bench both ways.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_call9
mkdir -p "$OUT"
step() { local name=$1 secs=$2; shift 2
    echo "== $name" | tee -a "$OUT/summary.txt"
    timeout "$secs" "$@" > "$OUT/$name.out" 2> "$OUT/$name.err"
    echo "   exit $?" | tee -a "$OUT/summary.txt"
    tail -n 3 "$OUT/$name.out" | cut -c1-400 | tee -a "$OUT/summary.txt"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
step 0_fuzz 300 python -m pytest tests/test_gpu_zx_neighbor_fuzz.py -q -m gpu -x
step 1_slab_tests 600 python -m pytest tests/test_gpu_slab_nccl.py tests/test_gpu_slab.py -q -m gpu
step 2_bench_p2p 400 $TR --master-port 29601 bench.py --gpus 2 --steps 30 --warmup 5
step 2_bench_nccl 400 $TR --master-port 29602 bench.py --gpus 2 --steps 30 --warmup 5 --nccl-exchange
step 3_bench_n1 300 python bench.py --steps 30 --warmup 5 --e2e-steps 4 --cpu-baseline-steps 0
cat "$OUT/summary.txt"
