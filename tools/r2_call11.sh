#!/bin/bash
# Round 2, GPU call 11 (2 GPUs): suite on HEAD (C host of the sharded channel, run_until chunks, quick reject in the
# general walk), phase times of the sharded step both ways, sharded bench both ways + through the C host, N = 1, ensemble.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_call11
mkdir -p "$OUT"
step() { local name=$1 secs=$2; shift 2
    echo "== $name" | tee -a "$OUT/summary.txt"
    timeout "$secs" "$@" > "$OUT/$name.out" 2> "$OUT/$name.err"
    echo "   exit $?" | tee -a "$OUT/summary.txt"
    tail -n 6 "$OUT/$name.out" | cut -c1-400 | tee -a "$OUT/summary.txt"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
export SPHB_ERRLOG=$PWD/$OUT/observed_errors.jsonl
step 1_tests 900 python -m pytest tests -q -m gpu
unset SPHB_ERRLOG
step 2_phases_p2p 300 $TR --master-port 29611 tools/slab_phases.py
step 2_phases_nccl 300 $TR --master-port 29612 tools/slab_phases.py --nccl-exchange
step 3_bench_p2p 400 $TR --master-port 29613 bench.py --gpus 2 --steps 30 --warmup 5
step 3_bench_nccl 400 $TR --master-port 29614 bench.py --gpus 2 --steps 30 --warmup 5 --nccl-exchange
step 3_bench_native 300 python tools/bench_native_slabs.py --slabs 2 --steps 200 --warmup 20
step 4_bench_n1 300 python bench.py --steps 30 --warmup 5 --e2e-steps 4 --cpu-baseline-steps 0
step 5_ens512 300 python bench.py --mode ensemble --channels 512 --steps 20 --warmup 3
cat "$OUT/summary.txt"
