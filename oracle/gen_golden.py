"""oracle/gen_golden.py -- TEST INFRASTRUCTURE.  Regenerates tests/golden/*.npz.

The reference ships no tests, golden vectors or fixtures, so the fixtures that pin this
repo's oracle are OUTPUTS OF THE REFERENCE ITSELF: its two MEX sources compiled unmodified
(oracle/_ref/libsph_ref.so, recipe oracle/Makefile) and driven by oracle/host.py with
OMP_NUM_THREADS=1 (deterministic pair-order sums).  Run in the authoring container, where
/root/reference exists:

    make -C oracle all && OMP_NUM_THREADS=1 python -m oracle.gen_golden

Fixtures (all fp64, column-major matrices stored as [n, k] arrays):
  ops_dp01_dev.npz      seeded developed state, dp=0.1 (540 particles): neighbour pairs + the outputs
                        of all eight physics modes on those inputs
  nbr_edge_cases.npz    neighbour search on inputs with unwrapped x (<0, >=DL), particles exactly on
                        the seam, coincident particles (r2 <= 1e-24 rejected) and a sliver last column
  traj_dp005.npz        dp=0.05 lattice start, the reference main-loop order for 120 steps (crosses the
                        cell sort at step 100): per-step dt / pair count / wall shear, state at 10 and 120
"""
from __future__ import annotations

import os
import sys

import numpy as np

os.environ.setdefault("OMP_NUM_THREADS", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import host as H  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

STATE_FIELDS = ("pos", "vel", "rho", "p", "drho_dt", "force", "force_prior", "Vol", "B")


def ops_case(be, prm, st):
    """Every MEX mode once, on one input state; returns {name: array}."""
    nf, nt = st.n_fluid, st.n_total
    out = {}
    pairs = be.neighbor_search(st.pos, nf, nt, prm.h, prm.DL)
    for k, name in enumerate(("pair_i", "pair_j", "dx", "dy", "r", "W", "dW")):
        out[name] = pairs[k]
    pi, pj, dx, dy, r, W, dW = pairs
    rho, Vol, B = be.physics("density_correction", pi, pj, dx, dy, r, W, dW, st.mass, nf, nt, prm.rho0, prm.h,
                             prm.inv_sigma0, nargout=3)
    out.update(dens_rho=rho, dens_Vol=Vol, dens_B=B)
    (fv,) = be.physics("viscous_force", pi, pj, dx, dy, r, dW, st.vel, Vol, B, prm.mu, prm.h, nf, nt, st.mass,
                       st.wall_vel, nargout=1)
    out["visc_force"] = fv
    fp = fv.copy(order="F")
    fp[:nf, 0] += st.mass[:nf] * prm.gravity_g
    (pt,) = be.physics("transport_correction", pi, pj, dx, dy, r, dW, Vol, B, st.pos, prm.h, nf, nt,
                       prm.transport_coeff, nargout=1)
    out["transp_pos"] = pt
    (pt_def,) = be.physics("transport_correction", pi, pj, dx, dy, r, dW, Vol, B, st.pos, prm.h, nf, nt, nargout=1)
    out["transp_pos_default_coeff"] = pt_def
    dt = H.verlet_time_step(st.vel[:nf], prm.c_f, prm.h, prm.nu, prm.gravity_g, np.inf)
    out["dt"] = np.array(dt)
    common = (pi, pj, dx, dy, r, dW, Vol, B, rho, st.mass, pt, st.vel, st.drho_dt, fp, dt, nf, nt, prm.rho0, prm.p0,
              prm.c_f, st.wall_vel)
    o = be.physics("integration_1st", *common, nargout=5)
    for k, name in enumerate(("rho", "p", "pos", "force", "drho_dt")):
        out["int1_" + name] = o[k]
    o2 = be.physics("integration_2nd", pi, pj, dx, dy, r, dW, Vol, o[0], o[2], st.vel, dt, nf, nt, st.wall_vel,
                    nargout=3)
    for k, name in enumerate(("pos", "drho_dt", "zeros")):
        out["int2_" + name] = o2[k]
    o = be.physics("integration_verlet", *common, nargout=6)
    for k, name in enumerate(("rho", "p", "pos", "vel", "drho_dt", "force")):
        out["verlet_" + name] = o[k]
    o = be.physics("advance_shell_step", pi, pj, dx, dy, r, W, dW, st.mass, st.pos, st.vel, st.wall_vel, st.rho,
                   st.drho_dt, dt, nf, nt, prm.rho0, prm.p0, prm.c_f, prm.mu, prm.h, prm.inv_sigma0, prm.gravity_g,
                   nargout=9)
    for k, name in enumerate(("rho", "p", "pos", "vel", "drho_dt", "force", "force_prior", "Vol", "B")):
        out["adv_" + name] = o[k]
    tb, tt = be.physics("wall_shear_monitor", pi, pj, dx, dy, r, dW, st.pos, st.vel, st.wall_vel, Vol, B, nf,
                        prm.DL, prm.DH, prm.mu, prm.h, nargout=2)
    out["tau"] = np.array([np.ravel(tb)[0], np.ravel(tt)[0]])
    return out


def edge_positions(prm):
    """Hand-made neighbour-search stress input (see module docstring)."""
    st = H.developed_state(prm, seed=777, jitter=0.25, noise=0.0)
    nf = st.n_fluid
    pos = st.pos.copy(order="F")
    pos[0, 0] = 0.0                      # exactly on the seam
    pos[1, 0] = prm.DL                   # == DL: wraps to 0
    pos[2, 0] = -1e-13                   # tiny negative: wraps to DL-eps
    pos[3, 0] = prm.DL - 1e-13
    pos[4, 0] = pos[5, 0]                # coincident pair (r2 = 0 <= 1e-24: rejected)
    pos[4, 1] = pos[5, 1]
    pos[6, 0] += 2.0 * prm.DL            # unwrapped input, two periods off
    pos[7, 0] -= 3.0 * prm.DL
    pos[8, 0] = 2.0 * prm.h              # exactly at the ghost threshold (strict <, nb:289)
    pos[9, 0] = prm.DL - 2.0 * prm.h     # exactly at the other threshold (strict >, nb:282)
    pos[10, 1] = pos[11, 1]
    pos[10, 0] = pos[11, 0] + 2.0 * prm.h  # pair exactly at the cut-off radius (strict <, nb:368)
    return pos, nf, st.n_total


def main():
    os.makedirs(OUT, exist_ok=True)
    be = H.RefBackend()

    # ---- ops_dp01_dev ------------------------------------------------------------------------
    prm = H.derive_params(dp=0.1, c_f=10.0)
    st = H.developed_state(prm, seed=12345)
    data = ops_case(be, prm, st)
    data.update(in_pos=st.pos, in_vel=st.vel, in_drho_dt=st.drho_dt, in_mass=st.mass, in_rho=st.rho,
                n_fluid=np.array(st.n_fluid), n_total=np.array(st.n_total), dp=np.array(0.1), c_f=np.array(10.0))
    np.savez_compressed(os.path.join(OUT, "ops_dp01_dev.npz"), **data)
    print("ops_dp01_dev: n_total", st.n_total, "pairs", data["pair_i"].size)

    # ---- nbr_edge_cases ----------------------------------------------------------------------
    edge = {}
    for tag, kw in (("a", dict(DL=1.0, DH=0.5, dp=0.05)), ("b", dict(DL=0.3, DH=0.4, dp=0.05)),
                    ("c", dict(DL=3.0, DH=1.0, dp=0.1))):
        prm_e = H.derive_params(c_f=10.0, **kw)
        pos, nf, nt = edge_positions(prm_e)
        pairs = be.neighbor_search(pos, nf, nt, prm_e.h, prm_e.DL)
        edge[f"{tag}_pos"] = pos
        edge[f"{tag}_meta"] = np.array([nf, nt, prm_e.h, prm_e.DL])
        for k, name in enumerate(("pair_i", "pair_j", "dx", "dy", "r", "W", "dW")):
            edge[f"{tag}_{name}"] = pairs[k]
        print("nbr_edge", tag, "n_total", nt, "pairs", pairs[0].size)
    np.savez_compressed(os.path.join(OUT, "nbr_edge_cases.npz"), **edge)

    # ---- traj_dp005 --------------------------------------------------------------------------
    prm = H.derive_params(dp=0.05, c_f=15.0, t_end=20.0, output_interval=1.0)
    st = H.init_particles(prm)
    H.initialise(be, st, prm)
    traj = {"n_pairs0": np.array(st.pairs[0].size)}
    dts, npairs, taus = [], [], []
    for k in range(1, 121):
        dts.append(H.reference_step(be, st, prm))
        npairs.append(st.pairs[0].size)
        taus.append((st.tau_bottom, st.tau_top))
        if k in (10, 120):
            for f in STATE_FIELDS:
                traj[f"s{k}_{f}"] = getattr(st, f).copy()
            traj[f"s{k}_t"] = np.array(st.t)
    traj.update(dt=np.array(dts), n_pairs=np.array(npairs), tau=np.array(taus))
    np.savez_compressed(os.path.join(OUT, "traj_dp005.npz"), **traj)
    print("traj_dp005: steps 120, t", st.t, "pairs", npairs[-1])


if __name__ == "__main__":
    main()
