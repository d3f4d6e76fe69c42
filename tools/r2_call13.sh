#!/bin/bash
# Round 2, GPU call 13 (2 GPUs; call 12 ran the same with three send launches): the send side as ONE launch; the exchange over the edge candidates only (3 small launches) and the receive side
# as one launch: slab parity (one GPU, two ranks over NCCL, C host on two GPUs), phase times, sharded bench three ways.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_call13
mkdir -p "$OUT"
step() { local name=$1 secs=$2; shift 2
    echo "== $name" | tee -a "$OUT/summary.txt"
    timeout "$secs" "$@" > "$OUT/$name.out" 2> "$OUT/$name.err"
    echo "   exit $?" | tee -a "$OUT/summary.txt"
    tail -n 6 "$OUT/$name.out" | cut -c1-400 | tee -a "$OUT/summary.txt"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
export SPHB_ERRLOG=$PWD/$OUT/observed_errors.jsonl
step 1_tests 900 python -m pytest tests/test_gpu_slab.py tests/test_gpu_slab_nccl.py tests/test_driver.py -q -m gpu
unset SPHB_ERRLOG
step 2_phases_p2p 300 $TR --master-port 29621 tools/slab_phases.py
step 2_phases_nccl 300 $TR --master-port 29622 tools/slab_phases.py --nccl-exchange
step 3_bench_p2p 400 $TR --master-port 29623 bench.py --gpus 2 --steps 30 --warmup 5
step 3_bench_nccl 400 $TR --master-port 29624 bench.py --gpus 2 --steps 30 --warmup 5 --nccl-exchange
step 3_bench_native 300 python tools/bench_native_slabs.py --slabs 2 --steps 200 --warmup 20
cat "$OUT/summary.txt"
