"""Runs the GPU test files against the CPU FUNCTIONAL SIMULATOR build of the same CUDA sources (tests/cpusim).

Test infrastructure, not parity evidence: the simulator executes the kernel bodies and the host logic of
libsph_b200 on the CPU (one ucontext fiber per CUDA thread, barriers at __syncthreads / *_sync, graphs as recorded
closures), which catches logic errors, bad indexing and broken host orchestration in code written without a GPU at
hand -- e.g. the batched ensemble launches and the flat staging build -- before GPU minutes are spent on them.  It
cannot see races, memory-model or alignment problems, and it measures nothing.  Parity is claimed from
`pytest -m gpu` on a B200 only.  The product never loads this library (SPHB_CPUSIM_LIB is read by tests/conftest.py).

The tests below run the quick part of the GPU suite.  The whole of it (everything but the two-rank NCCL test; about six
minutes) runs with

    B=tests/cpusim/_build; python tests/cpusim/build.py [-DSPHB_FLAT_STAGING=1]
    SPHB_CPUSIM_LIB=$PWD/$B/libsph_b200_cpusim.so SPHB_CPUSIM_DRIVER=$PWD/$B/sph_poiseuille_cpusim \
    SPHB_CPUSIM_MEXHOST=$PWD/$B/libsphb_mex_stubhost_cpusim.so python -m pytest tests -q -m gpu -k "not nccl"

Options: CPUSIM_SCHEDULE=reverse|random (order in which the threads of a CTA run between barriers), CPUSIM_ASAN=1 at
build time (AddressSanitizer; run with LD_PRELOAD=$(g++ -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0).
"""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

FAST = "not full_size and not ensemble_equals and not reference_loop and not golden_trajectory"


def _run(lib, files, k=None, env_extra=None, timeout=1500):
    env = dict(os.environ, SPHB_CPUSIM_LIB=lib, OMP_NUM_THREADS="1",
               SPHB_CPUSIM_DRIVER=os.path.join(os.path.dirname(lib), "sph_poiseuille_cpusim"),
               SPHB_CPUSIM_MEXHOST=os.path.join(os.path.dirname(lib), "libsphb_mex_stubhost_cpusim.so"),
               **(env_extra or {}))
    cmd = [sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", "-p", "no:cacheprovider"] + files
    if k:
        cmd += ["-k", k]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)
    tail = "\n".join((r.stdout + r.stderr).splitlines()[-25:])
    assert r.returncode == 0, tail
    return tail


@pytest.fixture(scope="module")
def simlib():
    sys.path.insert(0, os.path.join(ROOT, "tests", "cpusim"))
    import build as simbuild
    return simbuild.build()


def test_new_entry_points_on_the_simulator(simlib):
    """Batched ensemble launches (bit identity with the single-channel schedule, across cell sorts, n_inner 1 and 3)
    sphb_ctx_step_uploaded, and the seeded fuzz of the neighbour search (which found, on the simulator, the ghost
    entry that exact multiples of the cell size round one column further: GridDesc::ghost_anomaly)."""
    out = _run(simlib, ["tests/test_gpu_zz_batch.py", "tests/test_gpu_zy_step_uploaded.py",
                        "tests/test_gpu_zx_neighbor_fuzz.py", "tests/test_gpu_zw_step_fuzz.py"])
    assert " passed" in out and "skipped" not in out.splitlines()[-1], out


def test_standalone_driver_on_the_simulator(simlib):
    """config.ini -> 120 main-loop steps -> log lines, profile CSV, restart round trip (tests/test_driver.py) with the
    C driver linked against the simulator library."""
    out = _run(simlib, ["tests/test_driver.py"])
    assert " passed" in out, out


def test_slab_group_on_the_simulator(simlib):
    """Two x-slabs of one channel (pack / exchange / unpack / migration kernels of csrc/slab.cu, exchange buffers in
    CPU tensors) against the single-domain run: identical neighbour sets, fields within 1e-11.  The three-slab and
    the seam-migration cases of tests/test_gpu_slab.py pass too but take minutes on the simulator."""
    out = _run(simlib, ["tests/test_gpu_slab.py"], k="2-0.02 or halo_towards")
    assert " passed" in out, out


def test_c_host_of_the_sharded_channel_on_the_simulator(simlib):
    """sphb_slabs_* (csrc/slab_group.cu: cut, mailboxes, stepping, run_until, download -- no Python in the loop) against
    the single-domain run, three slabs sharing the simulator's one device; the same host drives one slab per GPU on a
    multi-GPU box (tests/test_gpu_slab.py::test_native_c_host_one_slab_per_gpu)."""
    out = _run(simlib, ["tests/test_gpu_slab.py"], k="native_c_host_matches and 3-0.02")
    assert " passed" in out, out


def test_parity_suite_subset_on_the_simulator(simlib):
    """The quick part of the device-resident and MEX-path parity tests (bit-exact neighbour lists, 1e-11 fields):
    guards the simulator itself and, with it, every later kernel change made without a GPU.  The fibers of a CTA are
    scheduled in a freshly shuffled order every round (CPUSIM_SCHEDULE=random): code that leans on an implicit
    order of threads between barriers would lose the bit-exact list parity here."""
    out = _run(simlib, ["tests/test_gpu_context.py", "tests/test_gpu_mex_path.py"], k=FAST,
               env_extra={"CPUSIM_SCHEDULE": "random"})
    assert " passed" in out, out
