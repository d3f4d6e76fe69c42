"""Batched ensemble launches (sphb_batch_*, csrc/batch.cu) against the same channels stepped one by one.

STATUS: written after the GPU budget of round 1 was spent.  The tests pass on the CPU functional simulator build of
the same sources (tests/cpusim, run by tests/test_cpusim.py; also under AddressSanitizer) -- kernel bodies, record
plumbing and host schedule are exercised there -- but have NOT yet run on a GPU, and no performance number is
claimed for the batched path.  The file sorts last so that a surprise on the hardware cannot mask the other tests.

The batch kernels call the very __device__ bodies of the single-channel kernels with per-channel arguments read
from device records, so the expected result is Channel.run_dual's, channel by channel (bit for bit on the simulator).
"""
import importlib

import numpy as np
import pytest

from parity_util import TOL, assert_close, log_observed

pytestmark = [pytest.mark.gpu]

FIELDS = ("pos", "vel", "rho", "p", "drho_dt", "force", "force_prior", "Vol", "B")


def _channels(sphb, n, sort_interval, DL=1.0, dp=0.05):
    cs = importlib.import_module("sph_b200.channel_setup")
    cfgs = [cs.ChannelConfig(DL=DL, dp=dp, mu=0.05 + 0.004 * k, U_bulk=0.4 + 0.015 * k, c_f=10.0).derive()
            for k in range(n)]
    pts = [cs.developed(c, seed=30 + k) for k, c in enumerate(cfgs)]

    def make():
        out = []
        for c, p in zip(cfgs, pts):
            ch = sphb.channel_from_config(c, p, sort_interval=sort_interval)
            ch.initialise()
            out.append(ch)
        return out

    return make


def _assert_same(a, b):
    """Batched launches run the same __device__ bodies as the single-channel kernels, but they are separate kernels
    to the compiler (arguments from a device record, other register budget), so the last bit of a sum may differ
    on the GPU (FMA contraction is decided per kernel); on the CPU simulator the two are bit-identical.  Fields are
    held to the field tolerance (measured on the B200, r2 call 10: at most 2.0e-13 over all batch tests, clock and
    monitors included), integers and flags exactly."""
    for k, (x, y) in enumerate(zip(a, b)):
        dx, dy = x.download(), y.download()
        for f in FIELDS:
            assert_close(dx[f], dy[f], f"channel {k} {f}", TOL)
        sx, sy = x.status(), y.status()
        for key in ("step", "n_half_pairs", "n_csr_entries"):
            assert sx[key] == sy[key], (k, key, sx[key], sy[key])
        for key in ("t", "dt", "vmax", "tau_bottom", "tau_top"):
            log_observed(f"channel {k} {key}", abs(sx[key] - sy[key]) / max(abs(sy[key]), 1e-300), 1e-12)
            assert abs(sx[key] - sy[key]) <= 1e-12 * max(abs(sy[key]), 1e-300), (k, key, sx[key], sy[key])
        assert sx["error_flags"] == 0


@pytest.mark.parametrize("n_inner", [1, 3])
def test_batch_equals_separate_runs(sphb, n_inner):
    """12 channels with different mu and g, 7 outer steps: batched launches == one channel after the other,
    including each channel's own clock, dt, max|v|, wall shear and pair counts."""
    make = _channels(sphb, 12, sort_interval=0)
    a, b = make(), make()
    batch = sphb.Batch(a)
    batch.run_dual(7, n_inner)
    batch.sync()
    for ch in b:
        ch.run_dual(7, n_inner)
    _assert_same(a, b)
    assert batch.error_flags() == 0
    assert batch.launches > 0
    batch.close(channels=True)
    for ch in b:
        ch.close()


def test_batch_across_cell_sorts(sphb):
    """sort_interval = 3: the per-channel cell sort swaps buffers in the middle of the batched schedule and the
    records are refreshed; 10 outer steps stay on the single-channel schedule."""
    make = _channels(sphb, 5, sort_interval=3)
    a, b = make(), make()
    batch = sphb.Batch(a)
    batch.run_dual(4, 1)
    batch.run_dual(6, 1)       # second call: graph already cached, sorts at outer steps 6 and 9
    batch.sync()
    for ch in b:
        ch.run_dual(10, 1)
    _assert_same(a, b)
    batch.close(channels=True)
    for ch in b:
        ch.close()


def test_batch_members_usable_after_destroy(sphb):
    """Destroying the batch hands the members their own streams back; they carry on alone."""
    make = _channels(sphb, 3, sort_interval=0)
    a, b = make(), make()
    batch = sphb.Batch(a)
    batch.run_dual(2, 1)
    batch.close()
    for ch in a:
        ch.run_dual(2, 1)
    for ch in b:
        ch.run_dual(4, 1)
    _assert_same(a, b)
    for ch in a + b:
        ch.close()


def test_batch_rejects_mixed_geometry(sphb):
    """Channels of different length cannot share a launch geometry: SPHB:batch."""
    one = _channels(sphb, 1, sort_interval=0, DL=1.0)()
    two = _channels(sphb, 1, sort_interval=0, DL=2.0)()
    with pytest.raises(sphb.SphError) as e:
        sphb.Batch(one + two)
    assert e.value.identifier == "SPHB:batch"
    for ch in one + two:
        ch.run_dual(1, 1)      # the failed creation left them on their own streams
        assert ch.status()["error_flags"] == 0
        ch.close()
