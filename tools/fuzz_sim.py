#!/usr/bin/env python
"""Seeded fuzz of libsph_b200 against the oracle beyond what the test suite runs, meant for the CPU functional
simulator (tests/cpusim) -- it also runs on a GPU, where it costs GPU minutes.  Test tooling, not product.

    python tests/cpusim/build.py
    SPHB_CPUSIM_LIB=$PWD/tests/cpusim/_build/libsph_b200_cpusim.so python tools/fuzz_sim.py <what> [cases] [seed]

  neighbors   random clouds of tests/test_gpu_zx_neighbor_fuzz.py with fresh seeds: seven pair columns bit-exact
  ulps        coordinates within 3 ulps of the branch points of the reference's arithmetic: bit-exact
  dense       clustered clouds (thousands of particles, hundreds per cell): stateless search bit-exact
  closure     resident context on the ulp clouds: half list exact, full rows = symmetric closure except pairs at
              r = 2h within rounding, never an entry the reference lacks
  steps       resident steps from adversarial states (violent velocities, moving walls, near-coincident particles)
  slab        slab group against the single-domain run (random length / slabs / drift / compaction cadence)
  batch       batched launches against separate runs (random channels, sort cadence, substeps, call pattern)

This is what found the ghost-entry corner cases of the neighbour list and the halo depth at the periodic seam.
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("OMP_NUM_THREADS", "1")

import conftest  # noqa: E402
from oracle import host as H  # noqa: E402
import test_gpu_zx_neighbor_fuzz as F  # noqa: E402

sphb = conftest.load_product()
be = H.best_backend()
FIELDS = ("pos", "vel", "rho", "p", "drho_dt", "force", "force_prior", "Vol", "B")


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    s = np.max(np.abs(b)) if b.size else 0.0
    return float(np.max(np.abs(a.reshape(b.shape) - b)) / (s if s > 0 else 1.0)) if b.size else 0.0


def same_pairs(got, want):
    return all(np.array_equal(a, b) for a, b in zip(got, want))


def ulp_clouds(seed, n_cases):
    rng = np.random.default_rng(seed)
    for case in range(n_cases):
        h = float(rng.choice([0.05, 0.1, 0.13, 0.0325, 0.065]))
        cell = 2.0 * h
        ncols = int(rng.integers(3, 12))
        DL = F._ulps(ncols * cell, int(rng.integers(-3, 4)))
        n = int(rng.integers(20, 200))
        nf = int(rng.integers(1, n))
        crit = [0.0, cell, DL - cell, DL, 2 * cell, DL - 2 * cell, (ncols - 1) * cell, (ncols - 2) * cell, 3 * cell,
                DL + cell, -cell, DL / 2]
        x = np.array([F._ulps(rng.choice(crit), int(rng.integers(-3, 4))) for _ in range(n)])
        k = rng.random(n) < 0.3
        x[k] = rng.random(int(k.sum())) * DL
        y = np.array([F._ulps(rng.choice([0.0, cell, 2 * cell, h, 0.5 * h]), int(rng.integers(-2, 3))) for _ in range(n)])
        k = rng.random(n) < 0.3
        y[k] = rng.random(int(k.sum())) * int(rng.integers(1, 5)) * cell
        yield f"ulps seed {seed} case {case}", np.asfortranarray(np.column_stack([x, y])), nf, n, h, DL


def fuzz_neighbors(n_cases, seed):
    bad = 0
    for block in range(1000 + seed, 1000 + seed + max(1, n_cases // 30)):
        for label, pos, nf, n, h, DL in F.cases(block):
            if not same_pairs(sphb.sph_neighbor_search_mex(pos, nf, n, h, DL), be.neighbor_search(pos, nf, n, h, DL)):
                bad += 1
                print("MISMATCH", label, flush=True)
    return bad


def fuzz_ulps(n_cases, seed):
    bad = 0
    for label, pos, nf, n, h, DL in ulp_clouds(seed, n_cases):
        if not same_pairs(sphb.sph_neighbor_search_mex(pos, nf, n, h, DL), be.neighbor_search(pos, nf, n, h, DL)):
            bad += 1
            print("MISMATCH", label, flush=True)
    return bad


def fuzz_dense(n_cases, seed):
    rng, bad = np.random.default_rng(seed), 0
    for case in range(n_cases):
        h = float(rng.choice([0.03, 0.05, 0.1]))
        cell = 2.0 * h
        DL = (int(rng.integers(2, 40)) - float(rng.choice([0.0, 0.5, 0.01]))) * cell
        n = int(rng.integers(500, 4000))
        nf = int(n * rng.uniform(0.5, 1.0))
        k, height = int(rng.integers(1, 8)), float(rng.integers(2, 10)) * cell
        centres = rng.random((k, 2)) * np.array([DL, height])
        pos = centres[rng.integers(0, k, n)] + rng.normal(0.0, float(rng.choice([0.2, 0.5, 1.5, 4.0])) * h, (n, 2))
        pos = np.asfortranarray(pos)
        if not same_pairs(sphb.sph_neighbor_search_mex(pos, nf, n, h, DL), be.neighbor_search(pos, nf, n, h, DL)):
            bad += 1
            print("MISMATCH dense case", case, n, nf, h, DL, flush=True)
    return bad


def _bare_channel(pos, nf, n, h, DL):
    dp = h / 1.3
    prm = sphb.ChannelParams(n_fluid=nf, n_total=n, DL=DL, DH=1.0, dp=dp, h=h, rho0=1.0, p0=100.0, c_f=10.0, mu=0.1,
                             gravity_g=0.8, inv_sigma0=dp * dp)
    ch = sphb.Channel(prm)
    z2, one = np.zeros((n, 2), order="F"), np.ones(n)
    ch.upload(pos=pos, vel=z2, rho=one, drho_dt=np.zeros(n), mass=one * dp * dp, wall_vel=z2)
    return ch


def fuzz_closure(n_cases, seed):
    bad = one_sided = 0
    for label, pos, nf, n, h, DL in ulp_clouds(seed, n_cases):
        want = be.neighbor_search(pos, nf, n, h, DL)
        ch = _bare_channel(pos, nf, n, h, DL)
        ch.neighbor_build()
        flags = ch.status()["error_flags"]
        got = ch.export_pairs()
        rp, col = ch.export_csr()
        ch.close()
        ok = flags == 0 and same_pairs(got[:2], want[:2])
        pi, pj = want[0].astype(int) - 1, want[1].astype(int) - 1
        rows = [[] for _ in range(nf)]
        r_of = {}
        for a, b, r in zip(pi, pj, want[4]):
            rows[a].append(b)
            r_of[(a, b)] = r_of[(b, a)] = r
            if b < nf:
                rows[b].append(a)
        for i in range(nf):
            w, g = set(rows[i]), set(col[rp[i]:rp[i + 1]].tolist())
            if g - w:
                ok = False           # an entry the reference lacks: never acceptable
            for j in w - g:
                one_sided += 1
                if abs(r_of[(i, j)] / (2.0 * h) - 1.0) > 1e-12:
                    ok = False       # a missing entry that is not at the cut-off
        if not ok:
            bad += 1
            print("MISMATCH", label, flush=True)
    print("one-sided pairs at r = 2h within rounding:", one_sided)
    return bad


def fuzz_steps(n_cases, seed):
    rng, bad = np.random.default_rng(seed), 0
    for case in range(n_cases):
        dp, DL = float(rng.choice([0.1, 0.08, 0.05])), float(rng.choice([0.4, 0.7, 1.0, 2.0]))
        prm = H.derive_params(DL=DL, DH=float(rng.choice([0.5, 1.0])), dp=dp, mu=float(rng.choice([0.01, 0.1, 1.0])),
                              c_f=float(rng.choice([5.0, 10.0, 20.0])), sort_interval=int(rng.choice([2, 100])))
        st = H.developed_state(prm, seed=int(rng.integers(1, 10 ** 6)), jitter=float(rng.choice([0.0, 0.2, 0.45])),
                               noise=float(rng.choice([0.0, 0.01, 0.3])))
        nf, mode = st.n_fluid, case % 4
        if mode == 1:
            st.vel[:nf] = rng.normal(0, 2.0, (nf, 2))
            st.drho_dt[:] = rng.normal(0, 50.0, st.n_total)
        elif mode == 2:
            st.wall_vel[nf:] = rng.normal(0, 1.0, (st.n_total - nf, 2))
            st.mass *= rng.uniform(0.5, 1.5, st.n_total)
        elif mode == 3:
            k = rng.integers(0, nf, 8)
            st.pos[k] = st.pos[(k + 1) % nf] + rng.normal(0, 1e-9, (8, 2))
        ref = st.copy()
        H.initialise(be, ref, prm)
        ch = sphb.channel_from_oracle_params(prm, st.n_fluid, st.n_total)
        ch.upload(pos=st.pos, vel=st.vel, rho=st.rho, drho_dt=st.drho_dt, mass=st.mass, wall_vel=st.wall_vel)
        ch.initialise()
        ok = True
        for k in range(3):
            dt = H.reference_step(be, ref, prm)
            ch.step()
            s, d = ch.status(), ch.download()
            worst = max(rel(d[f], getattr(ref, f)) for f in FIELDS)
            if s["n_half_pairs"] != ref.pairs[0].size or s["error_flags"] or abs(s["dt"] - dt) > 1e-11 * dt or not worst <= 1e-10:
                ok = False
                print("MISMATCH steps case", case, "mode", mode, "step", k, dp, DL, worst, flush=True)
                break
        ch.close()
        bad += not ok
    return bad


def fuzz_slab(n_cases, seed):
    cs = importlib.import_module("sph_b200.channel_setup")
    slab = importlib.import_module("sph_b200.slab")
    rng, bad = np.random.default_rng(seed), 0
    for case in range(n_cases):
        dp, DL = float(rng.choice([0.04, 0.05])), float(rng.choice([3.0, 3.7, 5.0, 6.02]))
        n_slabs = int(rng.choice([2, 3, 4]))
        cfg = cs.ChannelConfig(DL=DL, DH=1.0, dp=dp, c_f=10.0, sort_interval=0).derive()
        pt = cs.developed(cfg, seed=int(rng.integers(1, 10 ** 6)), jitter=float(rng.choice([0.1, 0.3])))
        pt.vel[:pt.n_fluid, 0] += float(rng.choice([0.0, 5.0, -8.0]))
        n_steps = int(rng.choice([6, 20]))
        try:
            grp = slab.SlabGroup(cfg, pt, n_slabs, compact_every=int(rng.choice([3, 7, 64])))
        except sphb.SphError:
            continue                 # slabs too narrow for their halo: refused loudly
        ch = sphb.channel_from_config(cfg, pt, sort_interval=0)
        ch.initialise()
        for _ in range(n_steps):
            ch.step()
            grp.step()
        want, got = ch.download(FIELDS[:5]), grp.gather(FIELDS[:5])
        nf = pt.n_fluid
        worst = max(rel(got[f], want[f][:nf]) for f in FIELDS[:5])
        ok = np.array_equal(got["ids"], np.arange(nf)) and grp.status()["error_flags"] == 0 and worst <= 1e-10
        if not ok:
            bad += 1
            print("MISMATCH slab case", case, dp, DL, n_slabs, n_steps, worst, flush=True)
        ch.close()
        grp.close()
    return bad


def fuzz_batch(n_cases, seed):
    cs = importlib.import_module("sph_b200.channel_setup")
    rng, bad = np.random.default_rng(seed), 0
    for case in range(n_cases):
        dp, DL = float(rng.choice([0.1, 0.0625, 0.05])), float(rng.choice([0.52, 1.0, 1.3, 2.0]))
        si, n_inner = int(rng.choice([0, 2, 3, 5])), int(rng.choice([1, 1, 2, 3]))
        cfgs = [cs.ChannelConfig(DL=DL, dp=dp, mu=float(rng.uniform(0.03, 0.3)), U_bulk=float(rng.uniform(0.2, 1.0)),
                                 c_f=float(rng.choice([8.0, 10.0, 15.0]))).derive() for _ in range(int(rng.integers(2, 7)))]
        pts = [cs.developed(c, seed=int(rng.integers(1, 10 ** 6)), jitter=0.2) for c in cfgs]

        def make():
            out = []
            for c, p in zip(cfgs, pts):
                ch = sphb.channel_from_config(c, p, sort_interval=si)
                ch.initialise()
                out.append(ch)
            return out

        a, b = make(), make()
        calls = [int(rng.integers(1, 6)) for _ in range(int(rng.integers(1, 4)))]
        batch = sphb.Batch(a)
        for n in calls:
            batch.run_dual(n, n_inner)
        batch.sync()
        ok = batch.error_flags() == 0
        for x, y in zip(a, b):
            for n in calls:
                y.run_dual(n, n_inner)
            dx, dy = x.download(), y.download()
            ok = ok and max(rel(dx[f], dy[f]) for f in FIELDS) <= 1e-10 and x.status()["step"] == y.status()["step"]
        batch.close(channels=True)
        for ch in b:
            ch.close()
        if not ok:
            bad += 1
            print("MISMATCH batch case", case, dp, DL, si, n_inner, calls, flush=True)
    return bad


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "neighbors"
    n_cases = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    seed = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    fn = dict(neighbors=fuzz_neighbors, ulps=fuzz_ulps, dense=fuzz_dense, closure=fuzz_closure, steps=fuzz_steps,
              slab=fuzz_slab, batch=fuzz_batch)[what]
    n_bad = fn(n_cases, seed)
    print(f"{what}: {n_cases} cases, seed {seed}: {n_bad} bad")
    sys.exit(1 if n_bad else 0)
