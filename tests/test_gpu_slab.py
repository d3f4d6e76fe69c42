"""GPU parity of the x-slab sharded channel against the single-domain device-resident run.

The reference has no domain decomposition, so the gate is "same answer as the single-domain run" (which
is itself pinned against the oracle in test_gpu_context.py): neighbour sets of every owned row identical
as sorted global index sets, fields within 1e-11 max-norm relative after several steps that include
migration across slab boundaries and across the periodic seam, identical time steps.
`SlabGroup` runs all slabs on ONE GPU with the same pack / unpack kernels and buffers the NCCL form uses.
"""
import numpy as np
import pytest

from parity_util import TOL, assert_close, log_observed

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _need_gpu(sphb):
    if not sphb.device_available():
        pytest.fail("no CUDA device: the product path has no CPU fallback")


def _setup(sphb, dp, DL, seed):
    import importlib
    cs = importlib.import_module("sph_b200.channel_setup")
    slab = importlib.import_module("sph_b200.slab")
    cfg = cs.ChannelConfig(DL=DL, DH=1.0, dp=dp, c_f=10.0, sort_interval=0).derive()
    pt = cs.developed(cfg, seed=seed, jitter=0.2)
    return cs, slab, cfg, pt


def _single(sphb, cfg, pt):
    ch = sphb.channel_from_config(cfg, pt, sort_interval=0)
    ch.initialise()
    return ch


def _global_rows(s, sl):
    """Neighbour sets of the owned rows of one slab as {global id: sorted global neighbour ids}."""
    n_owned, n_halo, n_wall = sl.counts()
    row_ptr, col = sl.ch.export_csr()                # rows of the whole capacity; those beyond the live rows are empty
    ids = np.zeros(n_owned + n_halo, dtype=np.int32)
    from ctypes import POINTER, c_int
    s._check(s.lib().sphb_ctx_slab_ids(sl.ch._ctx, ids.ctypes.data_as(POINTER(c_int))))
    gid = np.concatenate([ids.astype(np.int64), sl._keep["wall_ids"].astype(np.int64)])   # local row -> global id
    out = {}
    for r in range(n_owned):
        if ids[r] >= 0:                      # a row that migrated away (awaiting compaction) is nobody's
            out[int(ids[r])] = np.sort(gid[col[row_ptr[r]:row_ptr[r + 1]]])
    return out


@pytest.mark.parametrize("n_slabs,dp,DL,p2p", [(2, 0.02, 3.0, False), (3, 0.01, 2.0, False), (2, 0.02, 3.0, True),
                                               (4, 0.02, 4.0, True)])
def test_slab_group_matches_single_domain(sphb, n_slabs, dp, DL, p2p):
    """p2p: the peer-to-peer exchange (the pack kernel stores into the neighbours' mailboxes, the unpack kernel
    waits for its own; sphb_ctx_slab_p2p_*) instead of copies made by the caller -- the same records either way."""
    cs, slab, cfg, pt = _setup(sphb, dp, DL, seed=11)
    # a fast, seam-crossing flow so that rows migrate within a few steps
    ch = _single(sphb, cfg, pt)
    grp = slab.SlabGroup(cfg, pt, n_slabs, p2p=p2p)

    # neighbour sets of the initial list
    row_ptr, col = ch.export_csr()
    rows = {}
    for sl in grp.slabs:
        rows.update(_global_rows(sphb, sl))
    assert len(rows) == pt.n_fluid
    for i in range(pt.n_fluid):
        assert np.array_equal(rows[i], col[row_ptr[i]:row_ptr[i + 1]].astype(np.int64)), f"row {i}"

    owned0 = [sl.counts()[0] for sl in grp.slabs]
    n_steps = 12
    for k in range(n_steps):
        ch.step()
        grp.step()
        a, b = ch.status(), grp.status()
        assert b["error_flags"] == 0 and a["error_flags"] == 0
        # dt follows max|v|, and velocities agree to rounding only (the slabs sum in another order): measured 1.7e-15
        log_observed(f"dt at step {k}", abs(a["dt"] - b["dt"]) / a["dt"], 1e-13)
        assert abs(a["dt"] - b["dt"]) <= 1e-13 * a["dt"], f"dt differs at step {k}"
    want = ch.download(("pos", "vel", "rho", "p", "drho_dt"))
    got = grp.gather(("pos", "vel", "rho", "p", "drho_dt"))
    assert np.array_equal(got["ids"], np.arange(pt.n_fluid))
    nf = pt.n_fluid
    for f in ("pos", "vel", "rho", "p", "drho_dt"):
        assert_close(got[f], want[f][:nf], f, TOL)
    # the wall-shear monitor and max|v|: the slabs' partials reduced on the device (the SUM / MAX kernel of the NCCL
    # form) against the single-domain values of the same step
    sa = ch.status()
    vmax, tau_b, tau_t = grp.monitors()
    tau_scale = max(abs(sa["tau_bottom"]), abs(sa["tau_top"]))
    assert tau_scale > 0.0
    log_observed("tau_bottom", abs(tau_b - sa["tau_bottom"]) / tau_scale, TOL)
    log_observed("tau_top", abs(tau_t - sa["tau_top"]) / tau_scale, TOL)
    assert abs(tau_b - sa["tau_bottom"]) <= TOL * tau_scale, (tau_b, sa["tau_bottom"])
    assert abs(tau_t - sa["tau_top"]) <= TOL * tau_scale, (tau_t, sa["tau_top"])
    assert abs(vmax - sa["vmax"]) <= 1e-12 * sa["vmax"], (vmax, sa["vmax"])
    # final neighbour sets (after migration)
    row_ptr, col = ch.export_csr()
    rows = {}
    for sl in grp.slabs:
        rows.update(_global_rows(sphb, sl))
    assert len(rows) == pt.n_fluid
    bad = sum(0 if np.array_equal(rows[i], col[row_ptr[i]:row_ptr[i + 1]].astype(np.int64)) else 1
              for i in range(pt.n_fluid))
    assert bad == 0, f"{bad} rows differ after {n_steps} steps"
    assert sum(sl.owned(("rho",))["ids"].size for sl in grp.slabs) == pt.n_fluid
    ch.close()
    grp.close()
    _ = owned0


@pytest.mark.parametrize("n_slabs,dp,DL", [(3, 0.02, 3.0), (2, 0.01, 2.0)])
def test_native_c_host_matches_single_domain(sphb, n_slabs, dp, DL):
    """The sharded channel driven by the library's C host (sphb_slabs_*, csrc/slab_group.cu: cut, mailboxes, stepping --
    what `sph_poiseuille_b200 --slabs N` runs, no Python in the loop) against the single-domain run: steps with and
    without a remaining-time cut, then the reference loop up to a target (run_until), with a compaction in between."""
    cs, slab, cfg, pt = _setup(sphb, dp, DL, seed=21)
    pt.vel[:pt.n_fluid, 0] += 2.0            # rows migrate within the run
    ch = _single(sphb, cfg, pt)
    grp = slab.NativeSlabGroup(cfg, pt, n_slabs, devices=[0], compact_every=9)
    for k in range(14):
        remain = float("inf") if k % 5 else 0.6 * ch.status()["dt"]     # now and then the target cuts the step (m:259)
        ch.step(remain)
        grp.step(remain)
        a, b = ch.status(), grp.status()
        assert a["error_flags"] == 0 and b["error_flags"] == 0
        assert a["step"] == b["step"] and abs(a["t"] - b["t"]) <= 1e-13 * a["t"]
        assert abs(a["dt"] - b["dt"]) <= 1e-13 * a["dt"], f"dt differs at step {k}"
    t_target = ch.status()["t"] + 11.3 * ch.status()["dt"]
    n_a = ch.run_until(t_target, 1.0)
    n_b = grp.run_until(t_target, 1.0)
    assert n_a == n_b and n_a >= 11
    a, b = ch.status(), grp.status()
    assert abs(a["t"] - t_target) <= 1e-12 and abs(b["t"] - t_target) <= 1e-12
    tau_scale = max(abs(a["tau_bottom"]), abs(a["tau_top"]))
    assert abs(a["tau_bottom"] - b["tau_bottom"]) <= TOL * tau_scale and abs(a["tau_top"] - b["tau_top"]) <= TOL * tau_scale
    assert abs(a["vmax"] - b["vmax"]) <= 1e-12 * a["vmax"]
    want = ch.download(("pos", "vel", "rho", "p", "drho_dt"))
    got = grp.gather(("pos", "vel", "rho", "p", "drho_dt"))
    nf = pt.n_fluid
    for f in ("pos", "vel", "rho", "p", "drho_dt"):
        assert_close(got[f], want[f][:nf], f, TOL)
    ch.close()
    grp.close()


def test_native_c_host_one_slab_per_gpu(sphb):
    """One slab per GPU: every slab's step is ONE replayed CUDA graph on its own device, the slabs meet only in each
    other's mailboxes over NVLink (peer access inside one process).  Needs two GPUs (the driver's 1-GPU box skips it;
    run by the multi-GPU calls of the round, profiles/README.md)."""
    import torch
    n_dev = torch.cuda.device_count()
    if n_dev < 2:
        pytest.skip("needs at least two GPUs")
    n_slabs = min(n_dev, 4)
    cs, slab, cfg, pt = _setup(sphb, 0.01, 4.0, seed=31)
    pt.vel[:pt.n_fluid, 0] += 2.0
    ch = _single(sphb, cfg, pt)
    grp = slab.NativeSlabGroup(cfg, pt, n_slabs, compact_every=25)
    assert sorted(grp.member_device(r) for r in range(n_slabs)) == list(range(n_slabs))
    ch.run_dual(60, 1)
    grp.run(60)                      # graph replays, two compaction steps in between
    grp.sync()
    a, b = ch.status(), grp.status()
    assert a["error_flags"] == 0 and b["error_flags"] == 0 and a["step"] == b["step"] == 60
    assert abs(a["t"] - b["t"]) <= 1e-12 * a["t"]
    want = ch.download(("pos", "vel", "rho", "p", "drho_dt"))
    got = grp.gather(("pos", "vel", "rho", "p", "drho_dt"))
    for f in ("pos", "vel", "rho", "p", "drho_dt"):
        assert_close(got[f], want[f][:pt.n_fluid], f, TOL)
    t_target = a["t"] + 17.5 * a["dt"]
    assert ch.run_until(t_target, 1.0) == grp.run_until(t_target, 1.0)
    assert abs(grp.status()["t"] - t_target) <= 1e-12
    ch.close()
    grp.close()


def test_slab_rows_migrate_across_the_seam(sphb):
    """Uniform fast drift: within a few hundred steps rows cross slab boundaries and x = DL; the slabs
    keep every particle exactly once and stay equal to the single-domain run."""
    cs, slab, cfg, pt = _setup(sphb, 0.02, 2.0, seed=5)
    pt.vel[:pt.n_fluid, 0] += 3.0           # strong mean flow: ~0.45 cell columns per 100 steps
    ch = _single(sphb, cfg, pt)
    grp = slab.SlabGroup(cfg, pt, 2, compact_every=40, p2p=True)     # several compactions inside the run
    start = [set(sl.owned(("rho",))["ids"].tolist()) for sl in grp.slabs]
    for _ in range(150):
        ch.step()
        grp.step()
    end = [set(sl.owned(("rho",))["ids"].tolist()) for sl in grp.slabs]
    moved = sum(len(a - b) for a, b in zip(start, end))
    assert moved > 0, "no row migrated: the test does not exercise migration"
    assert sorted(i for s in end for i in s) == list(range(pt.n_fluid))
    want = ch.download(("pos", "vel", "rho"))
    got = grp.gather(("pos", "vel", "rho"))
    nf = pt.n_fluid
    # 150 steps of independent rounding (different summation orders); measured on the B200: 2.1e-14 (r2 call 10)
    for f in ("pos", "vel", "rho"):
        assert_close(got[f], want[f][:nf], f, TOL)
    assert grp.status()["error_flags"] == 0
    ch.close()
    grp.close()


def test_halo_towards_the_seam_covers_four_cells_with_a_sliver_last_column(sphb):
    """dp = 0.05, DL = 3: 23.08 cell columns.  Counted in columns, the copies the slab of column 0 receives from
    across the periodic seam would span 3.08 cells; a step consumes four, and rows that a fast drift carries over
    x = 0 came out 1e-6 off in drho/dt (1e-8 in rho after 6 steps).  With four full cells of copies the slabs
    stay on the single-domain run."""
    cs, slab, cfg, pt = _setup(sphb, 0.05, 3.0, seed=493)
    pt.vel[:pt.n_fluid, 0] -= 8.0
    ch = _single(sphb, cfg, pt)
    grp = slab.SlabGroup(cfg, pt, 2)
    for _ in range(8):
        ch.step()
        grp.step()
    assert grp.status()["error_flags"] == 0
    want = ch.download(("pos", "vel", "rho", "p", "drho_dt"))
    got = grp.gather(("pos", "vel", "rho", "p", "drho_dt"))
    nf = pt.n_fluid
    assert np.array_equal(got["ids"], np.arange(nf))
    for f in ("pos", "vel", "rho", "p", "drho_dt"):
        assert_close(got[f], want[f][:nf], f, TOL)   # measured 1.2e-13 (r2 call 10)
    ch.close()
    grp.close()


def test_slabs_at_size_with_the_sliver_of_the_dl96_channel(sphb):
    """BASELINE.json configs[4] geometry in small: dp = 0.0025 and a channel length whose last cell column is a
    0.23-cell sliver (DL = 6.0075: 924.23 columns; DL = 96 has 14769.23), two slabs of ~490 000 rows each against the
    single-domain device run of the same 980 000-particle state.  A strong drift carries rows over both slab cuts
    and over x = 0 within the steps taken."""
    cs, slab, cfg, pt = _setup(sphb, 0.0025, 6.0075, seed=96)
    ncols = cfg.DL / (2.0 * cfg.h)
    assert 0.2 < ncols - np.floor(ncols) < 0.26 and pt.n_total > 950_000
    pt.vel[:pt.n_fluid, 0] -= 20.0          # dt is the viscous limit 1.3e-5 here: ~1.2 dp of drift in 12 steps
    ch = _single(sphb, cfg, pt)
    grp = slab.SlabGroup(cfg, pt, 2)
    start = [set(sl.owned(("rho",))["ids"].tolist()) for sl in grp.slabs]
    for k in range(12):
        ch.step()
        grp.step()
        a, b = ch.status(), grp.status()
        assert b["error_flags"] == 0 and a["error_flags"] == 0
        assert abs(a["dt"] - b["dt"]) <= 1e-13 * a["dt"], f"dt differs at step {k}"
    end = [set(sl.owned(("rho",))["ids"].tolist()) for sl in grp.slabs]
    assert sum(len(x - y) for x, y in zip(start, end)) > 0, "no row migrated"
    want = ch.download(("pos", "vel", "rho", "p", "drho_dt"))
    got = grp.gather(("pos", "vel", "rho", "p", "drho_dt"))
    nf = pt.n_fluid
    assert np.array_equal(got["ids"], np.arange(nf))
    for f in ("pos", "vel", "rho", "p", "drho_dt"):
        assert_close(got[f], want[f][:nf], f, TOL)
    sa = ch.status()
    vmax, tau_b, tau_t = grp.monitors()
    tau_scale = max(abs(sa["tau_bottom"]), abs(sa["tau_top"]))
    assert abs(tau_b - sa["tau_bottom"]) <= TOL * tau_scale and abs(tau_t - sa["tau_top"]) <= TOL * tau_scale
    # neighbour sets of the rows next to the periodic seam and next to the slab cut (the full comparison of a
    # million rows is the small cases' job): identical as sorted global index sets
    row_ptr, col = ch.export_csr()
    rows = {}
    for sl in grp.slabs:
        rows.update(_global_rows(sphb, sl))
    assert len(rows) == nf
    x = want["pos"][:nf, 0]
    cell = 2.0 * cfg.h
    cut = 0.5 * cfg.DL
    near = np.nonzero((x < 5 * cell) | (x > cfg.DL - 5 * cell) | (np.abs(x - cut) < 5 * cell))[0]
    assert near.size > 10_000
    bad = sum(0 if np.array_equal(rows[int(i)], col[row_ptr[i]:row_ptr[i + 1]].astype(np.int64)) else 1 for i in near)
    assert bad == 0, f"{bad} of {near.size} rows near the seam / the cut differ"
    ch.close()
    grp.close()
