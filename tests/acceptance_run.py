#!/usr/bin/env python
"""Acceptance statistics of a full run (BASELINE.json north_star: "after the full run, the
centreline / edge u_x profile within 1e-3 U_max of the reference and L2 vs the analytic u(y) no worse than
the reference's").  Not a pytest module: a checker script, run by hand or by gpurun.

    python tests/acceptance_run.py --impl b200      --dp 0.04 --c_f 10 --end 5  --out profiles/acc_b200.json
    python tests/acceptance_run.py --impl reference --dp 0.04 --c_f 10 --end 5  --out profiles/acc_ref.json
    python tests/acceptance_run.py --compare profiles/acc_b200.json profiles/acc_ref.json

Both arms run the reference's single-rate main loop (SPH_Poiseuille.m:246-302) from the lattice at rest to
end_time with output points every `--out-interval`; `b200` through the device-resident context of
libsph_b200, `reference` through the reference's own C code (oracle/_ref; the pinned port when absent)
driven by oracle/host.py.  The reference is not run-to-run reproducible with > 1 OpenMP thread and its
flow is noisy, so profiles are compared as x- and time-averages over the last output
points, and the spread of those averages is reported beside every difference.

Statistics (definitions: SURVEY.md 8(c)): full-length binned u(y) and its L2 vs analytic
(SPH_Poiseuille_postprocess.m:42); slice profiles at x0 = DL/2 (centre) and x0 = 0 (seam) with
|x - x0| <= max(dp, h) periodic (m:592-605), n_bins = max(20, round(DH / dp)); edge gap =
max |u_seam - u_centre| / U_max; edge L2 = the postprocess formula on the seam profile; u_y RMS / U_max.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def binned(y, u, lo, hi, n_bins):
    edges = np.linspace(lo, hi, n_bins + 1)
    mid = 0.5 * (edges[:-1] + edges[1:])
    b = np.searchsorted(edges, y, side="right") - 1
    b[y == edges[-1]] = n_bins - 1
    ok = (y >= edges[0]) & (y <= edges[-1])
    s = np.bincount(b[ok], weights=u[ok], minlength=n_bins)[:n_bins]
    c = np.bincount(b[ok], minlength=n_bins)[:n_bins]
    m = s / np.maximum(c, 1)
    m[c == 0] = np.nan
    return mid, m


def slice_profile(pos, ux, DL, DH, x0, half, n_bins):
    xw = np.mod(pos[:, 0], DL)
    d = np.abs(xw - x0)
    d = np.minimum(d, DL - d)
    sel = d <= half
    return binned(pos[sel, 1], ux[sel], 0.0, DH, n_bins)[1]


def l2(u, exact):
    ok = ~np.isnan(u)
    return math.sqrt(np.sum((u[ok] - exact[ok]) ** 2) / max(np.sum(exact[ok] ** 2), np.finfo(float).eps))


def stats_of(pos_f, vel_f, P):
    n_bins = max(20, int(math.floor(P["DH"] / P["dp"] + 0.5)))
    y_mid, full = binned(pos_f[:, 1], vel_f[:, 0], 0.0, P["DH"], n_bins)
    half = max(P["dp"], P["h"])
    centre = slice_profile(pos_f, vel_f[:, 0], P["DL"], P["DH"], 0.5 * P["DL"], half, n_bins)
    seam = slice_profile(pos_f, vel_f[:, 0], P["DL"], P["DH"], 0.0, half, n_bins)
    exact = P["g"] / (2.0 * P["nu"]) * y_mid * (P["DH"] - y_mid)
    return dict(y=y_mid.tolist(), full=full.tolist(), centre=centre.tolist(), seam=seam.tolist(),
                L2_full=l2(full, exact), L2_edge=l2(seam, exact),
                edge_gap=float(np.nanmax(np.abs(seam - centre)) / P["U_max"]),
                uy_rms=float(np.sqrt(np.mean(vel_f[:, 1] ** 2)) / P["U_max"]),
                ux_max=float(np.max(vel_f[:, 0])))


def run_b200(a):
    from conftest import load_product
    import importlib
    sphb = load_product()
    cs = importlib.import_module("sph_b200.channel_setup")
    cfg = cs.ChannelConfig(DL=a.DL, DH=a.DH, dp=a.dp, c_f=a.c_f, mu=a.mu, U_bulk=a.U_bulk, end_time=a.end,
                           output_interval=a.out_interval, sort_interval=a.sort_interval,
                           transport_coeff=a.transport).derive()
    pt = cs.lattice(cfg)
    ch = sphb.channel_from_config(cfg, pt)
    ch.initialise()
    P = dict(DL=cfg.DL, DH=cfg.DH, dp=cfg.dp, h=cfg.h, g=cfg.gravity_g, nu=cfg.nu, U_max=cfg.U_max)
    outs, steps, t0 = [], 0, time.perf_counter()
    t_next = cfg.output_interval
    while t_next <= cfg.end_time + 1e-12:
        steps += ch.run_until(t_next, cfg.end_time)
        d = ch.download(("pos", "vel"))
        s = stats_of(d["pos"][:pt.n_fluid], d["vel"][:pt.n_fluid], P)
        st = ch.status()
        s.update(t=st["t"], tau_bottom=st["tau_bottom"], tau_top=st["tau_top"])
        outs.append(s)
        t_next += cfg.output_interval
    wall = time.perf_counter() - t0
    ch.close()
    return dict(impl="b200", params=P, n_fluid=pt.n_fluid, n_total=pt.n_total, steps=steps, wall_s=wall, outputs=outs)


def run_reference(a):
    os.environ.setdefault("OMP_NUM_THREADS", str(a.threads))
    from oracle import host as H
    be = H.best_backend()
    prm = H.derive_params(DL=a.DL, DH=a.DH, dp=a.dp, mu=a.mu, U_bulk=a.U_bulk, c_f=a.c_f, sort_interval=a.sort_interval,
                          transport_coeff=a.transport)
    st = H.init_particles(prm)
    H.initialise(be, st, prm)
    P = dict(DL=prm.DL, DH=prm.DH, dp=prm.dp, h=prm.h, g=prm.gravity_g, nu=prm.nu, U_max=1.5 * a.U_bulk)
    outs, steps, t0 = [], 0, time.perf_counter()
    t_next = a.out_interval
    nf = st.n_fluid
    while t_next <= a.end + 1e-12:
        while st.t < t_next - 1e-12:
            H.reference_step(be, st, prm, remain=min(t_next - st.t, a.end - st.t))
            steps += 1
        s = stats_of(st.pos[:nf], st.vel[:nf], P)
        s.update(t=st.t, tau_bottom=st.tau_bottom, tau_top=st.tau_top)
        outs.append(s)
        t_next += a.out_interval
    return dict(impl="reference", kind=be.kind, threads=int(os.environ["OMP_NUM_THREADS"]), params=P, n_fluid=nf,
                n_total=st.n_total, steps=steps, wall_s=time.perf_counter() - t0, outputs=outs)


def averaged(run, last):
    o = run["outputs"][-last:]
    avg = {k: np.nanmean(np.array([x[k] for x in o], dtype=float), axis=0) for k in ("full", "centre", "seam")}
    spread = {k: np.nanstd(np.array([x[k] for x in o], dtype=float), axis=0) for k in ("full", "centre", "seam")}
    return avg, spread, {k: float(np.mean([x[k] for x in o])) for k in ("L2_full", "L2_edge", "edge_gap", "uy_rms")}


def compare(fa, fb, last):
    A, B = json.load(open(fa)), json.load(open(fb))
    U = A["params"]["U_max"]
    a_avg, a_sp, a_sc = averaged(A, last)
    b_avg, b_sp, b_sc = averaged(B, last)
    res = {"a": {"impl": A["impl"], "steps": A["steps"], "wall_s": A["wall_s"], **a_sc},
           "b": {"impl": B["impl"], "steps": B["steps"], "wall_s": B["wall_s"], **b_sc}, "last_outputs_averaged": last}
    for k in ("full", "centre", "seam"):
        res[f"max_diff_{k}_over_Umax"] = float(np.nanmax(np.abs(a_avg[k] - b_avg[k])) / U)
        res[f"spread_{k}_over_Umax"] = {"a": float(np.nanmax(a_sp[k]) / U), "b": float(np.nanmax(b_sp[k]) / U)}
    res["final_L2_full"] = {"a": A["outputs"][-1]["L2_full"], "b": B["outputs"][-1]["L2_full"]}
    print(json.dumps(res, indent=1))
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", choices=("b200", "reference"))
    ap.add_argument("--compare", nargs=2)
    ap.add_argument("--last", type=int, default=5, help="output points averaged in --compare")
    ap.add_argument("--DL", type=float, default=3.0)
    ap.add_argument("--DH", type=float, default=1.0)
    ap.add_argument("--dp", type=float, default=0.04)
    ap.add_argument("--mu", type=float, default=0.1)
    ap.add_argument("--U_bulk", type=float, default=0.666667)
    ap.add_argument("--c_f", type=float, default=10.0)
    ap.add_argument("--end", type=float, default=5.0)
    ap.add_argument("--out-interval", type=float, default=0.2)
    ap.add_argument("--sort-interval", type=int, default=100)
    ap.add_argument("--transport", type=float, default=0.30)
    ap.add_argument("--threads", type=int, default=1)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    if a.compare:
        res = compare(a.compare[0], a.compare[1], a.last)
        if a.out:
            json.dump(res, open(a.out, "w"), indent=1)
        return 0
    run = run_b200(a) if a.impl == "b200" else run_reference(a)
    o = run["outputs"][-1]
    print(f"{run['impl']}: {run['steps']} steps in {run['wall_s']:.1f} s, t = {o['t']:.4f}, L2_full = {100 * o['L2_full']:.3f} %, "
          f"L2_edge = {100 * o['L2_edge']:.3f} %, edge gap = {o['edge_gap']:.4f}, u_y rms = {o['uy_rms']:.4f}")
    if a.out:
        json.dump(run, open(a.out, "w"))
    return 0


if __name__ == "__main__":
    sys.exit(main())
