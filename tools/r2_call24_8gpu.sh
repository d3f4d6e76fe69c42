#!/bin/bash
# Round 2, GPU call 24 (8 GPUs; call 14 with the cheaper fences of the exchange kernels, the walk changes and tiles-per-CTA of the batched ensemble): BASELINE configs[4] (DL = 96 sharded over 8 GPUs) and configs[3] (4096-channel
# ensemble, 512 per GPU) on HEAD, weak scaling N = 4 / 8, the C host on 8 GPUs, multi-GPU parity tests.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_call24
mkdir -p "$OUT"
step() { local name=$1 secs=$2; shift 2
    echo "== $name" | tee -a "$OUT/summary.txt"
    timeout "$secs" "$@" > "$OUT/$name.out" 2> "$OUT/$name.err"
    echo "   exit $?" | tee -a "$OUT/summary.txt"
    tail -n 6 "$OUT/$name.out" | cut -c1-330 | tee -a "$OUT/summary.txt"; }
TR() { echo python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2; }
nvidia-smi -L | tee "$OUT/gpus.txt"
step 1_tests 600 python -m pytest tests/test_gpu_slab.py tests/test_gpu_slab_nccl.py -q -m gpu -k "one_slab_per_gpu or nccl"
step 2_scale_n8 400 $(TR 8 29631) bench.py --gpus 8 --steps 40 --warmup 5
step 2_scale_n4 400 $(TR 4 29632) bench.py --gpus 4 --steps 40 --warmup 5
step 3_dl96_n8 600 $(TR 8 29633) bench.py --gpus 8 --DL 12 --steps 30 --warmup 5 --e2e-steps 3
step 4_ens4096_n8 600 $(TR 8 29634) bench.py --gpus 8 --mode ensemble --channels 512 --steps 20 --warmup 3
step 5_native_n8 400 python tools/bench_native_slabs.py --slabs 8 --steps 200 --warmup 20
step 5_native_dl96 600 python tools/bench_native_slabs.py --slabs 8 --DL 12 --steps 60 --warmup 10
step 7_scale_n2 400 $(TR 2 29636) bench.py --gpus 2 --steps 40 --warmup 5
step 6_phases_n8 300 $(TR 8 29635) tools/slab_phases.py
cat "$OUT/summary.txt"
