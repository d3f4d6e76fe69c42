#!/bin/bash
# First GPU call of round 2: verifies and measures what round 1 wrote after its GPU budget was spent.
#
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/r2_first_gpu_call.sh'
#
# Everything lands in gpurun_out/r2_first/ (one log / JSON line per step); a failing step does not stop the rest,
# and every step runs under its own `timeout` so that an unverified path cannot hang the box.
#   1. the default build: GPU parity tests + the N = 1 bench line (baseline of the round)
#   2. batched ensemble launches (csrc/batch.cu): parity tests, then per-channel graphs vs --batched
#   3. flat staging (SPHB_FLAT_STAGING=1 build): GPU parity tests + bench line, 2 and 4 slots per pass, then the
#      default build is restored
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2_first
mkdir -p "$OUT"
step() {   # step <name> <seconds> <command...>
    local name=$1 secs=$2
    shift 2
    echo "== $name" | tee -a "$OUT/summary.txt"
    timeout "$secs" "$@" > "$OUT/$name.out" 2> "$OUT/$name.err"
    echo "   exit $?" | tee -a "$OUT/summary.txt"
    tail -n 3 "$OUT/$name.out" | cut -c1-400 | tee -a "$OUT/summary.txt"
}

step 1_tests_default 900 python -m pytest tests -x -q -m gpu
step 1_bench_default 600 python bench.py --steps 20 --warmup 3

step 2_batch_tests 600 python -m pytest tests/test_gpu_zz_batch.py -x -q -m gpu
if grep -q " passed" "$OUT/2_batch_tests.out" && ! grep -q "failed\|error" "$OUT/2_batch_tests.out"; then
    step 2_ens256_graphs 600 python bench.py --mode ensemble --channels 256 --steps 20 --warmup 3
    step 2_ens256_batched 600 python bench.py --mode ensemble --channels 256 --steps 20 --warmup 3 --batched
    step 2_ens512_batched 900 python bench.py --mode ensemble --channels 512 --steps 20 --warmup 3 --batched
else
    echo "   batched ensemble tests not green: benches skipped" | tee -a "$OUT/summary.txt"
fi

step 3_build_flat 900 env SPHB_FLAT_STAGING=1 python sph-poiseuille-flow_b200/build.py
if [ -f sph-poiseuille-flow_b200/libsph_b200.so ]; then
    step 3_tests_flat 900 python -m pytest tests -x -q -m gpu
    step 3_bench_flat 600 python bench.py --steps 20 --warmup 3
fi
for u in 2 4; do   # slots per thread and staging pass (default 3)
    step 3_build_flat_u$u 900 env SPHB_FLAT_STAGING=1 SPHB_STAGE_UNROLL=$u python sph-poiseuille-flow_b200/build.py
    step 3_bench_flat_u$u 600 python bench.py --steps 20 --warmup 3 --e2e-steps 2 --cpu-baseline-steps 0
done
step 3_build_default 900 python sph-poiseuille-flow_b200/build.py
step 3_tests_default_again 900 python -m pytest tests -x -q -m gpu
cat "$OUT/summary.txt"
