#!/usr/bin/env python
"""bench.py -- particle-substeps/s of the SPH particle loop (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--inner I] [--impl b200|reference]

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): the dp = 0.0025,
DL = 3, DH = 1 periodic channel (480 000 fluid + 9 600 wall particles), seeded synthetic developed
state.  One bench "step" = one OUTER step of the particle loop: kernel-sum density + KGC, viscous
(+ body) force, transport-velocity shift, `--inner` acoustic Verlet substeps (device dt, immediate
periodic wrap), [cell sort when due], neighbour-list rebuild, wall-shear monitor.  The default
--inner 1 is exactly one pass of the reference main loop (SPH_Poiseuille.m:250-283), the only
schedule the reference has at HEAD and -- at the reference's mu = 0.1 and dp = 0.0025, where the
viscous limit 0.125 h^2/nu is below the acoustic one -- also what a dual-criteria scheduler
degenerates to (freezing the viscous force over several substeps is unstable there).  A
particle-substep is one integration_verlet application to one particle; value = n_total * inner *
steps * ranks / time, so all outer-step work (list rebuild included) is amortised into the number.

Multi-GPU (launched by torchrun, one rank per GPU), "scaling": "weak" (fixed work per GPU):
  --mode slab (default for N > 1): ONE periodic channel of length N * DL cut into x-slabs on cell-column
      boundaries (BASELINE.json configs[4] in weak-scaling form): every rank steps its slab + 4 halo
      columns, and once per step ONE uneven NCCL all-to-all moves the migrating rows and halo copies to the
      two neighbour ranks and max|v| / the shear partials to everybody; the whole step (library kernels +
      the collective) is replayed as one CUDA graph (--no-graph for eager launches);
  --mode replicas: N independent channels, one per rank, no data-path collective;
  --mode ensemble [--channels C]: C independent dp = 0.02 channels PER GPU with mu and g swept on a grid
      (BASELINE.json configs[3]; 4096 in all at 8 GPUs with C = 512), stepped together by
      sphb_ensemble_run_dual; no communication.  With --batched the channels are stepped by batched
      launches instead (sphb_batch_run_dual: every kernel launched once for all channels).
Timing: CUDA events on the library's stream, barrier + synchronize on both sides, max over ranks.

Lines printed: ONE JSON line on rank 0 (see the task contract: value, e2e, roofline, cpu_baseline,
clocks, gpu_launches).  `--impl reference` times the reference's own OpenMP C code (oracle/_ref,
built unmodified from the reference sources; the plain-C port when that is absent) on the host cores
with the same schedule.  The oracle is used ONLY for that arm and for the cpu_baseline leg.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "particle-substeps/s"
ALG_BYTES_SUBSTEP = 184           # P6 contract, 112 B read + 72 B written per particle
# per-particle algorithmic bytes of each kernel (inputs read once + outputs written once, fp64)
ALG_BYTES_KERNEL = {
    "force_fused": 112 + 80,       # P2+P3+dt+P4: pos vel rho drho Vol B mass wall_vel -> F_prior force v_new pos (rho,p)half
    "pressure": 112 + 48,          # pos vel rho drho F_prior Vol B mass -> force, v_new, pos
    "continuity": 56 + 56,         # v_new Vol rho drho pos -> rho p drho pos vel
    "density": 24 + 16,            # snapshot pos, mass -> rho, Vol
    "kgc": 24 + 32,                # snapshot pos, Vol -> B
    "visc_transport": 88 + 32,     # vel Vol B pos wall_vel -> force_prior, pos
    "nbr_walk_count": 16 + 8,      # snapshot pos -> two counts
    "nbr_walk_fill": 16 + 86,      # pos -> ~20.5 x 4 B CSR + row pointer
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--inner", type=int, default=1, help="acoustic substeps per outer step (1 = reference loop)")
    ap.add_argument("--dual-auto", type=int, default=0, metavar="MAX_INNER",
                    help="N = 1: step with the device-side dual-criteria scheduler (sphb_ctx_run_dual_auto: outer Dt and "
                         "acoustic dt decided on the device, at most MAX_INNER substeps per outer step); the value counts "
                         "the substeps actually taken.  With --mu below ~0.02 an outer step takes several.")
    ap.add_argument("--mu", type=float, default=None, help="dynamic viscosity (default: the reference's 0.1)")
    ap.add_argument("--impl", default="b200", choices=("b200", "reference"))
    ap.add_argument("--dp", type=float, default=0.0025)
    ap.add_argument("--DL", type=float, default=3.0)
    ap.add_argument("--no-flush", action="store_true", help="do not flush L2 between timed steps")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--profile-steps", type=int, default=5)
    ap.add_argument("--cpu-baseline-steps", type=int, default=20, help="outer steps of the CPU sample (0 = skip)")
    ap.add_argument("--seed", type=int, default=12345)
    ap.add_argument("--no-presort", action="store_true",
                    help="time from the raw lattice order instead of the order after the reference's first cell "
                         "sort (m:272, 529-568)")
    ap.add_argument("--no-graph", action="store_true", help="sharded mode: eager launches instead of one CUDA graph per step")
    ap.add_argument("--nccl-exchange", action="store_true",
                    help="sharded mode: move the records with one NCCL all-to-all per step instead of the peer-to-peer "
                         "mailboxes (the pack kernel storing into the neighbours' memory over NVLink; the default)")
    ap.add_argument("--channels", type=int, default=512,
                    help="--mode ensemble: independent channels PER GPU (BASELINE configs[3]: 4096 in all at 8 GPUs)")
    ap.add_argument("--batched", action="store_true", default=True,
                    help="--mode ensemble: step the channels with batched launches (sphb_batch_*: every kernel "
                         "launched once for all channels; the default, measured 8.8e8 vs 7.2e8 particle-substeps/s)")
    ap.add_argument("--per-channel-graphs", dest="batched", action="store_false",
                    help="--mode ensemble: one context + stream + CUDA graph per channel instead of batched launches")
    ap.add_argument("--mode", default="auto", choices=("auto", "slab", "replicas", "ensemble"),
                    help="N > 1: 'slab' = ONE periodic channel of length N*DL cut into x-slabs (halo exchange + "
                         "migration over NCCL, weak scaling); 'replicas' = N independent channels (ensemble); "
                         "auto = slab")
    return ap.parse_args()


def load_product():
    from __graft_entry__ import load_product as lp
    return lp()


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


NCU_EXTRA = {}      # pipe utilisation of the kernel ncu_traffic() was last asked about (same capture)


def ncu_traffic(kernel: str):
    """DRAM bytes per launch of `kernel` from the newest committed ncu summary (profiles/*_ncu_summary.json,
    written by tools/ncu_to_profile.py from an `ncu --set full` capture of this same command)."""
    import glob
    names = {"force_fused": "k_force_fused", "pressure": "k_pressure", "continuity": "k_continuity",
             "density": "k_density", "kgc": "k_kgc", "visc_transport": "k_visc_transport", "nbr_walk": "k_walk_tile"}
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_ncu_summary.json")))
    if not files or kernel not in names:
        return None, None
    try:
        rec = json.load(open(files[-1])).get(names[kernel])
        if not rec:
            return None, None
        NCU_EXTRA.clear()
        NCU_EXTRA.update({k: rec[0].get(k) for k in ("duration_us", "fp64_pipe_pct", "lsu_wavefronts_pct",
                                                     "issue_active_pct", "warps_active_pct", "dram_throughput_pct",
                                                     "fp64_thread_instructions")})
        NCU_EXTRA["capture"] = os.path.relpath(files[-1], ROOT)
        return rec[0]["dram_bytes"], os.path.relpath(files[-1], ROOT)
    except (OSError, ValueError, KeyError, IndexError):
        return None, None


def workload_name(cfg, pt, inner):
    return (f"dp={cfg.dp:g} DL={cfg.DL:g} DH={cfg.DH:g} channel, {pt.n_fluid}+{pt.n_total - pt.n_fluid} particles, "
            f"{inner} acoustic substeps per outer step")


# ------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md: sample DURING the timed region)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.idx = device_index
        self.proc = None
        self.lines = []
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# CPU reference (oracle/_ref or the port) on the same state and schedule
# ------------------------------------------------------------------------------------------------
def cpu_run(cfg, pt, inner, n_steps, n_warm):
    """Times n_steps outer steps of the dual schedule through the CPU oracle.  Returns (particle-
    substeps/s, kind, threads, seconds)."""
    threads = host_threads()
    os.environ["OMP_NUM_THREADS"] = str(threads)   # before libgomp loads
    from oracle import host as H
    if not (os.path.exists(H.REF_SO) or os.path.exists(H.PORT_SO)):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "all"], check=True, capture_output=True)
    be = H.best_backend()
    prm = H.derive_params(DL=cfg.DL, DH=cfg.DH, dp=cfg.dp, rho0=cfg.rho0, mu=cfg.mu, U_bulk=cfg.U_bulk, c_f=cfg.c_f,
                          sort_interval=cfg.sort_interval, transport_coeff=cfg.transport_coeff)
    st = H.init_particles(prm)
    for f in ("pos", "vel", "rho", "drho_dt", "mass", "wall_vel"):
        setattr(st, f, getattr(pt, f).copy(order="F"))
    H.initialise(be, st, prm)
    for _ in range(n_warm):
        H.dual_outer_step(be, st, prm, inner)
    t0 = time.perf_counter()
    for _ in range(n_steps):
        H.dual_outer_step(be, st, prm, inner)
    dt = time.perf_counter() - t0
    return st.n_total * inner * n_steps / dt, be.kind, threads, dt


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    lp = load_product()
    sharded = args.gpus > 1 and args.mode in ("auto", "slab")      # the B200 arm then runs ONE channel of length N * DL
    cfg = lp.ChannelConfig(DL=args.DL * (args.gpus if sharded else 1), dp=args.dp).derive()
    # bounded sample: shrink the channel length (same dp, same physics, per-particle cost unchanged)
    # until (steps + warmup) outer steps fit in ~150 s at ~1.2e-6 s per particle-substep-equivalent
    DL = cfg.DL
    est = lambda L: (args.steps + args.warmup) * (L / cfg.dp) * (cfg.DH / cfg.dp + 8) * (2.0e-6 + 0.35e-6 * args.inner)  # noqa: E731
    while est(DL) > 150.0 and DL > 0.3:
        DL *= 0.5
    cfg_s = lp.ChannelConfig(DL=DL, dp=args.dp).derive()
    pt = lp.channel_setup.developed(cfg_s, seed=args.seed)
    if not args.no_presort:
        pt = lp.channel_setup.cell_sorted(cfg_s, pt)
    value, kind, threads, secs = cpu_run(cfg_s, pt, args.inner, args.steps, args.warmup)
    sample = (f"{args.steps} outer steps x {args.inner} substeps of the dp={cfg_s.dp:g} channel "
              f"{'at full length' if DL == cfg.DL else f'shortened to DL={cfg_s.DL:g}'} ({pt.n_total} particles)")
    full = lp.channel_setup.lattice(cfg)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "particle-substeps/s", "n_gpus": args.gpus,
        # ms per step OF THE LABELLED WORKLOAD at the measured per-particle rate (= the measured time when the sample is
        # the full channel; extrapolated when the channel was shortened to bound the run: cpu_baseline.sample says which)
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * full.n_total * args.inner / max(value, 1e-300),
        "ms_per_step_sample": 1e3 * secs / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(cfg, full, args.inner), "inner_substeps": args.inner,
                   "list_rebuild": "every outer step", "parallelism": "OpenMP host threads",
                   "measured_on": ("the labelled channel" if DL == cfg.DL else
                                   f"the same channel shortened to DL={cfg_s.DL:g} ({pt.n_total} particles); per-particle rate")},
        "cpu_baseline": {"value": value, "unit": "particle-substeps/s", "cores": threads, "kind": kind,
                         "sample": sample},
        "e2e": {"value": value, "unit": "particle-substeps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
# the B200 arm
# ------------------------------------------------------------------------------------------------
def b200_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (libsph_b200 has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    lp = load_product()
    if args.mode == "ensemble":
        return b200_ensemble_arm(args, lp, rank, local_rank, world, barrier, max_over_ranks, sum_over_ranks)
    if world > 1 and args.mode in ("auto", "slab"):
        return b200_slab_arm(args, lp, rank, local_rank, world, barrier, max_over_ranks, sum_over_ranks)
    cfg = (lp.ChannelConfig(DL=args.DL, dp=args.dp) if args.mu is None else
           lp.ChannelConfig(DL=args.DL, dp=args.dp, mu=args.mu)).derive()
    pt = lp.channel_setup.developed(cfg, seed=args.seed + rank)     # one independent channel per rank
    if not args.no_presort:
        pt = lp.channel_setup.cell_sorted(cfg, pt)   # the row order of a run past its first sort point (m:272)
    nt, inner, K, W = pt.n_total, args.inner, args.steps, args.warmup
    ch = lp.channel_from_config(cfg, pt, device=local_rank)
    ch.initialise()
    stream = torch.cuda.ExternalStream(ch.stream, device=torch.device("cuda", local_rank))
    flush = None if args.no_flush else torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    # ---- device-resident throughput ----------------------------------------------------------------
    auto = args.dual_auto if world == 1 else 0
    if auto:
        inner = 1          # (the kernel profile and the e2e leg below run the reference schedule)
        ch.run_dual_auto(max(W, 3), auto)
        c0 = ch.dual_counts()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            ev0.record(stream)
            ch.run_dual_auto(K, auto)                   # K outer steps, one CUDA graph each; returns when they are done
            ev1.record(stream)
        ch.sync()
        c1 = ch.dual_counts()
        auto_ms = ev0.elapsed_time(ev1)
        auto_line = {"outer_steps": c1["outer_steps"] - c0["outer_steps"], "sub_steps": c1["sub_steps"] - c0["sub_steps"],
                     "max_inner": auto, "last_Dt": c1["last_Dt"], "ms_per_outer_step": auto_ms / K,
                     "value": nt * (c1["sub_steps"] - c0["sub_steps"]) / (auto_ms * 1e-3),
                     "l2": "no flush: consecutive outer steps of one simulation"}
    ch.run_dual(max(W, 3), inner)
    ch.sync()
    l0 = ch.timing()["launches"]
    step0 = int(ch.status()["step"]) // max(inner, 1)   # outer steps so far: which timed steps carry a cell sort (m:272)
    si = int(cfg.sort_interval)
    sorts_timed = sum(1 for k in range(step0 + 1, step0 + K + 1) if si > 0 and k % si == 0 and k != 1)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    with torch.cuda.stream(stream):
        for k in range(K):
            if flush is not None:
                flush.zero_()                     # > L2 (126 MB): the next step starts from HBM
            ev[k][0].record(stream)
            ch.run_dual(1, inner)
            ev[k][1].record(stream)
    ch.sync()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = max_over_ranks(sum(a.elapsed_time(b) for a, b in ev))
    launches = ch.timing()["launches"] - l0
    st = ch.status()
    if st["error_flags"] or not np.isfinite(st["vmax"]) or st["vmax"] > 10 * cfg.U_max:
        raise SystemExit(f"bench.py: run invalid (flags={st['error_flags']}, vmax={st['vmax']})")
    value = world * nt * inner * K / (ms_total * 1e-3)

    # ---- end to end through the C ABI with host buffers --------------------------------------------
    Ke = max(1, min(args.e2e_steps, K))

    def pinned(cols):
        t = torch.empty((cols, nt) if cols > 1 else (nt,), dtype=torch.float64).pin_memory().numpy()
        return t.T if cols > 1 else t                                  # [nt, cols] column-major view
    in_fields = (("pos", 2), ("vel", 2), ("rho", 1), ("drho_dt", 1))
    out_fields = (("pos", 2), ("vel", 2), ("rho", 1), ("p", 1), ("drho_dt", 1), ("force", 2))
    host = {k: pinned(c) for k, c in in_fields}                        # pinned inputs of a step
    outs = {k: pinned(c) for k, c in out_fields}                       # pinned results of a step
    ch.download_into(host)

    # One pass of m:251-283 for a caller whose state lives on the host: sphb_ctx_step_uploaded = list for the
    # uploaded positions + shear monitor + physics (no list rebuild for positions the next upload replaces).
    # inner > 1 (or a library without that entry point) takes the longer route through run_dual.
    e2e_path = {"uploaded": inner == 1 and hasattr(ch, "step_uploaded")}

    e2e_path["one_call"] = e2e_path["uploaded"] and hasattr(ch, "step_host")

    def e2e_step():
        if e2e_path["one_call"]:      # the same three calls as ONE (the list build overlaps the rest of the upload)
            ch.step_host(host, outs)
            for k in host:
                host[k], outs[k] = outs[k], host[k]
            return outs
        ch.upload(pos=host["pos"], vel=host["vel"], rho=host["rho"], drho_dt=host["drho_dt"])
        if e2e_path["uploaded"]:
            try:
                ch.step_uploaded()
            except lp.SphError as err:
                print(f"bench.py: sphb_ctx_step_uploaded failed ({err}); e2e continues through run_dual", file=sys.stderr)
                e2e_path["uploaded"] = False
        if not e2e_path["uploaded"]:
            ch.neighbor_build()
            ch.run_dual(1, inner)
        ch.download_into(outs)
        for k in host:                                                 # the results are the next step's inputs
            host[k], outs[k] = outs[k], host[k]
        return outs

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(Ke):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_value = world * nt * inner * Ke / e2e_s
    if e2e_path["uploaded"]:
        ch.neighbor_build()                                            # a fresh list for the kernel profile below
    h2d = nt * (16 + 16 + 8 + 8)
    d2h = nt * (16 + 16 + 8 + 8 + 8 + 16)

    # ---- per-kernel device time, CUDA events on the launching stream (eager launches) --------------
    roof = None
    kernels = {}
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except (OSError, ValueError):
            pass
        hbm = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        ch.kernel_profile(enable=1, reset=True)
        ch.run_dual(max(1, args.profile_steps), inner)
        prof = ch.kernel_profile(enable=0, reset=True)
        total = sum(ms for ms, _ in prof.values()) or 1.0
        for name, (ms, calls) in prof.items():
            if calls:
                kernels[name] = {"ms_per_call": ms / calls, "calls": calls, "share": ms / total}
        top = max((n for n in kernels if n in ALG_BYTES_KERNEL), key=lambda n: kernels[n]["share"])
        per_launch_bytes = ALG_BYTES_KERNEL[top] * nt
        achieved = per_launch_bytes / (kernels[top]["ms_per_call"] * 1e-3) / 1e9
        traffic, traffic_src = ncu_traffic(top)
        roof = {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": hbm, "unit": "GB/s",
                "frac": achieved / hbm, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "alg_bytes_per_launch": per_launch_bytes, "ncu": dict(NCU_EXTRA),
                "alg_bytes_per_particle": ALG_BYTES_KERNEL[top], "ms_per_launch": kernels[top]["ms_per_call"],
                "share_of_step": kernels[top]["share"],
                "step_frac": (value / world) * ALG_BYTES_SUBSTEP / (hbm * 1e9),
                "fp64_note": "fp64 gather over ~20 neighbours per particle: the FP64 pipe and the shared-memory "
                             "data pipe bound these kernels, not HBM (profiles/README.md)"}
        # the kernels of the acoustic substep alone (forces + continuity), list rebuild and KGC excluded
        # FP64 roofline of the same kernel: FP64-pipe thread instructions per launch (counted once per profile by
        # ncu, smsp__thread_inst_executed_pipe_fp64 of the committed capture named in roofline.ncu.capture) over the
        # live launch time, against the DFMA rate measured on this device just now (sphb_fp64_peak)
        try:
            peak_tf = lp.fp64_peak()
        except Exception as err:   # noqa: BLE001
            peak_tf = None
            print(f"bench.py: fp64 peak not measured ({err})", file=sys.stderr)
        n_inst = NCU_EXTRA.get("fp64_thread_instructions")
        if peak_tf:
            fp64 = {"peak_tflops": peak_tf, "peak_source": "sphb_fp64_peak: DFMA chains, this device, this run",
                    "peak_ginst_s": peak_tf * 1e3 / 2.0}
            if n_inst:
                ginst = n_inst / (kernels[top]["ms_per_call"] * 1e-3) / 1e9
                fp64.update({"achieved_ginst_s": ginst, "achieved_tflops_dfma_equiv": 2.0 * ginst / 1e3,
                             "frac": ginst / (peak_tf * 1e3 / 2.0), "inst_per_particle": n_inst / nt,
                             "inst_source": NCU_EXTRA.get("capture")})
            roof["fp64"] = fp64
        sub_ms = sum(kernels[n]["ms_per_call"] for n in ("force_fused", "time_step", "pressure", "continuity", "clock")
                     if n in kernels)
        roof["substep_kernels_only"] = {"ms": sub_ms, "value": nt / (sub_ms * 1e-3),
                                        "frac": nt / (sub_ms * 1e-3) * ALG_BYTES_SUBSTEP / (hbm * 1e9)}

    # ---- CPU baseline beside it (rank 0, N = 1 only) ------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and args.cpu_baseline_steps > 0:
        v, kind, threads, secs = cpu_run(cfg, pt, inner, args.cpu_baseline_steps, 1)
        cpu = {"value": v, "unit": "particle-substeps/s", "cores": threads, "kind": kind,
               "sample": f"{args.cpu_baseline_steps} outer steps x {inner} substeps of the same {nt}-particle state "
                         f"({secs:.1f} s, 1 warm-up step)"}
    total_launches = sum_over_ranks(float(launches))
    ch.close()
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "particle-substeps/s", "n_gpus": world, "steps": K, "warmup": max(W, 3),
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(cfg, pt, inner), "inner_substeps": inner,
                       "list_rebuild": "every outer step (device counting sort + tile-local 16-bit lists)",
                       "cell_sort_interval": cfg.sort_interval,
                       "cell_sorts_in_timed_steps": sorts_timed,
                       "row_order": ("raw lattice" if args.no_presort else
                                     "after the reference's cell sort (m:529-568), as in a run past its first sort point"),
                       "l2": ("256 MiB memset between timed steps" if flush is not None else
                              "no flush: consecutive steps of one simulation"),
                       "parallelism": f"{world} independent channel(s), one per GPU"},
            "e2e": {"value": e2e_value, "unit": "particle-substeps/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": Ke,
                    "path": ("sphb_ctx_step_host: upload -> list + shear + physics -> download in one call, pinned host "
                             "(the list build overlaps the rest of the upload)" if e2e_path.get("one_call") else
                             "sphb_ctx_upload -> sphb_ctx_step_uploaded (list + shear + physics) -> sphb_ctx_download, "
                             "pinned host" if e2e_path["uploaded"] else
                             "sphb_ctx_upload -> neighbor_build -> run_dual(1, inner) -> sphb_ctx_download, pinned host")},
            "gpu_launches": int(total_launches),
            "roofline": roof, "kernels": kernels, "cpu_baseline": cpu, "clocks": clocks,
            "final_state": {"vmax": st["vmax"], "n_half_pairs": st["n_half_pairs"], "t": st["t"]},
        }
        if auto:
            line["dual_auto"] = auto_line     # device-side dual-criteria scheduler: particle-substeps/s over the same state
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0



# ------------------------------------------------------------------------------------------------
# N > 1: one periodic channel of length N * DL, x-slab sharded
# ------------------------------------------------------------------------------------------------
def b200_slab_arm(args, lp, rank, local_rank, world, barrier, max_over_ranks, sum_over_ranks):
    import importlib

    import numpy as np
    import torch
    import torch.distributed as dist

    slabmod = importlib.import_module("sph_b200.slab")
    cfg = lp.ChannelConfig(DL=args.DL * world, dp=args.dp, sort_interval=0).derive()
    pt = lp.channel_setup.developed(cfg, seed=args.seed)              # the same global state on every rank
    if not args.no_presort:
        pt = lp.channel_setup.cell_sorted(cfg, pt)                    # row order after the reference's cell sort
    K, W = args.steps, max(args.warmup, 3)
    dev = torch.device("cuda", local_rank)
    sc = slabmod.SlabChannel(cfg, pt, rank, world, dev, use_graph=not args.no_graph, p2p=not args.nccl_exchange)
    ch = sc.slab.ch
    stream = sc.stream                           # the library's kernels, torch and NCCL share this stream in slab mode
    flush = None if args.no_flush else torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    # the first warm-up step runs eagerly and counts the kernels of one step (a replayed graph launches the
    # same kernels but does not pass through the library's launch counter)
    want_graph, sc.use_graph = sc.use_graph, False
    l_eager = ch.timing()["launches"]
    sc.step()
    launches_per_step = ch.timing()["launches"] - l_eager
    sc.use_graph = want_graph
    for _ in range(W - 1):
        sc.step()
    ch.sync()
    l0 = ch.timing()["launches"]
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    for k in range(K):
        with torch.cuda.stream(stream):
            if flush is not None:
                flush.zero_()                     # > L2 (126 MB): the next step starts from HBM
            ev[k][0].record(stream)
        sc.step()                                 # physics | pack -> NCCL send/recv + all-gather -> unpack | list + shear
        ev[k][1].record(stream)
    ch.sync()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = max_over_ranks(sum(a.elapsed_time(b) for a, b in ev))
    launches = ch.timing()["launches"] - l0
    st = ch.status()
    flags = max_over_ranks(float(st["error_flags"]))
    if flags or not np.isfinite(st["vmax"]) or st["vmax"] > 10 * cfg.U_max:
        raise SystemExit(f"bench.py: sharded run invalid (flags={flags}, vmax={st['vmax']})")
    value = pt.n_total * K / (ms_total * 1e-3)
    n_owned, n_halo, n_wall = sc.slab.counts()      # owned region (incl. rows awaiting compaction), halo, walls

    graph_used = bool(sc.use_graph and sc._graph_ok and sc._graph is not None)
    if graph_used:
        launches = launches_per_step * K      # every replay launches the kernels of one (captured) step
    # ---- end to end: the rank's rows live in PINNED host memory between steps: upload -> halo exchange -> list ->
    # physics -> download, every step (sphb_ctx_slab_rows_io: pos, vel, rho, drho_dt of the owned region)
    Ke = max(1, min(args.e2e_steps, K))
    n_rows = n_owned                                   # owned region incl. rows awaiting compaction (they stay dead)

    def pinned(cols):
        t = torch.empty((cols, n_rows) if cols > 1 else (n_rows,), dtype=torch.float64).pin_memory().numpy()
        return t.T if cols > 1 else t
    host = {"pos": pinned(2), "vel": pinned(2), "rho": pinned(1), "drho_dt": pinned(1)}
    sc.slab.rows_io(True, n_rows, host)
    ch.sync()

    def e2e_step():
        with torch.cuda.stream(stream):
            sc.slab.rows_io(False, n_rows, host)                 # H2D of the rank's rows
            sc.slab.step_begin()                                 # density .. continuity
            sc.exchange()                                        # migrants + halo copies between the ranks
            sc.slab.step_end()                                   # list + shear
            sc.slab.rows_io(True, n_rows, host)                  # D2H of the results
        ch.sync()

    # (rows that migrate in during these steps are appended beyond n_rows and stay on the device: a few per step)
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(Ke):
        e2e_step()
    torch.cuda.synchronize()
    ch.sync()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_value = pt.n_total * Ke / e2e_s
    nl = n_owned + n_halo + n_wall
    h2d = int(sum_over_ranks(float(n_rows * 48)))
    d2h = int(sum_over_ranks(float(n_rows * 48)))

    roof, kernels = None, {}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    sc.use_graph = False                       # per-kernel events need eager launches
    ch.kernel_profile(enable=1, reset=True)
    for _ in range(max(1, args.profile_steps)):
        sc.step()
    prof = ch.kernel_profile(enable=0, reset=True)
    if rank == 0:
        total = sum(ms for ms, _ in prof.values()) or 1.0
        for name, (ms, calls) in prof.items():
            if calls:
                kernels[name] = {"ms_per_call": ms / calls, "calls": calls, "share": ms / total}
        top = max((n for n in kernels if n in ALG_BYTES_KERNEL), key=lambda n: kernels[n]["share"])
        achieved = ALG_BYTES_KERNEL[top] * nl / (kernels[top]["ms_per_call"] * 1e-3) / 1e9
        traffic, traffic_src = ncu_traffic(top)
        roof = {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": hbm, "unit": "GB/s",
                "frac": achieved / hbm, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "alg_bytes_per_particle": ALG_BYTES_KERNEL[top], "ms_per_launch": kernels[top]["ms_per_call"],
                "share_of_step": kernels[top]["share"], "rows_per_launch": nl, "ncu": dict(NCU_EXTRA),
                "step_frac": (value / world) * ALG_BYTES_SUBSTEP / (hbm * 1e9),
                "kernel_ms_per_step_rank0": total / max(1, args.profile_steps)}
    total_launches = sum_over_ranks(float(launches))
    sc.close()
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "particle-substeps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(cfg, pt, 1) + f", x-slab sharded over {world} GPUs "
                                   f"({args.DL:g} of length per GPU)", "inner_substeps": 1,
                       "list_rebuild": "every step (device counting sort + tile-local lists)",
                       "cell_sort_interval": 0,
                       "l2": ("256 MiB memset between timed steps" if flush is not None else
                              "no flush: consecutive steps of one simulation"),
                       "parallelism": f"x-slabs on cell-column boundaries, {slabmod.HALO_COLS} halo columns, one "
                                      f"exchange per step: " + ("migration + halo copies, max|v| and shear partials in one "
                                      "uneven NCCL all-to-all" if args.nccl_exchange else "the pack kernel stores migrants, halo "
                                      "copies, max|v| and shear partials straight into the other ranks' mailboxes over NVLink "
                                      "(CUDA IPC peer memory, sequence-number arrival flags; no collective)"),
                       "step_launch": ("one CUDA graph per step (library kernels" + (" + NCCL all-to-all)" if args.nccl_exchange
                                                                                  else " only)")) if graph_used else "eager",
                       "rows_rank0": {"owned": n_owned, "halo": n_halo, "wall": n_wall}},
            "e2e": {"value": e2e_value, "unit": "particle-substeps/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": Ke,
                    "path": "per rank, pinned host rows: sphb_ctx_slab_rows_io (H2D) -> exchange -> list -> physics -> "
                            "sphb_ctx_slab_rows_io (D2H)"},
            "gpu_launches": int(total_launches),
            "roofline": roof, "kernels": kernels, "cpu_baseline": None, "clocks": clocks,
            "final_state": {"vmax": st["vmax"], "t": st["t"], "step": st["step"]},
        }
        print(json.dumps(line), flush=True)
    dist.destroy_process_group()
    return 0


# ------------------------------------------------------------------------------------------------
# ensembles of independent small channels with swept mu / g (BASELINE.json configs[3])
# ------------------------------------------------------------------------------------------------
def b200_ensemble_arm(args, lp, rank, local_rank, world, barrier, max_over_ranks, sum_over_ranks):
    import numpy as np
    import torch

    dp = args.dp if args.dp != 0.0025 else 0.02            # configs[3] is quoted on dp = 0.02 channels
    n_ch, K, W = args.channels, args.steps, max(args.warmup, 3)
    total = n_ch * world
    side = int(np.ceil(np.sqrt(total)))
    chans, n_part = [], 0
    for k in range(n_ch):
        g_idx = rank * n_ch + k                            # this channel's place on the (mu, g) grid
        mu = 0.05 * (0.2 / 0.05) ** ((g_idx % side) / max(side - 1, 1))          # log-spaced 0.05 .. 0.2
        ub = 0.333333 * (4.0) ** ((g_idx // side) / max(side - 1, 1))            # g spans a factor 4 through U_bulk
        cfg = lp.ChannelConfig(DL=args.DL, dp=dp, mu=mu, U_bulk=ub, c_f=10.0 * max(1.0, ub / 0.666667)).derive()
        pt = lp.channel_setup.developed(cfg, seed=args.seed + g_idx)
        pt = lp.channel_setup.cell_sorted(cfg, pt)
        ch = lp.channel_from_config(cfg, pt, device=local_rank)
        ch.initialise()
        chans.append(ch)
        n_part += pt.n_total
    ens = lp.Batch(chans) if args.batched else lp.Ensemble(chans)
    ens.run_dual(W, 1)
    ens.sync()
    l0 = ens.launches if args.batched else sum(ch.timing()["launches"] for ch in chans)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    t0 = time.perf_counter()
    ens.run_dual(K, 1)
    ens.sync()
    secs = max_over_ranks(time.perf_counter() - t0)      # many streams: wall clock around a synchronised region
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = (ens.launches if args.batched else sum(ch.timing()["launches"] for ch in chans)) - l0
    bad = sum(1 for ch in chans if ch.status()["error_flags"])
    n_all = int(sum_over_ranks(float(n_part)))
    value = n_all * K / secs
    total_launches = sum_over_ranks(float(launches))
    if args.batched:
        ens.close(channels=True)
    else:
        ens.close()
    if bad:
        raise SystemExit(f"bench.py: {bad} channels of the ensemble raised device flags")
    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": "particle-substeps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": 1e3 * secs / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"ensemble of {total} independent dp={dp:g} DL={args.DL:g} channels "
                                   f"({n_all} particles), mu and g swept on a {side} x {side} grid", "inner_substeps": 1,
                       "parallelism": (f"{n_ch} channels per GPU, batched launches (grid = tiles x channels, one CUDA "
                                       f"graph per step for all of them), no communication" if args.batched else
                                       f"{n_ch} channels per GPU, one context + stream + CUDA graph each, no communication"),
                       "l2": "working set of the ensemble exceeds L2; no flush", "timing": "host clock around a "
                       "device-synchronised region (one stream per channel)"},
            "e2e": None, "gpu_launches": int(total_launches), "roofline": None, "cpu_baseline": None, "clocks": clocks}),
            flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    return b200_arm(args)


if __name__ == "__main__":
    sys.exit(main())
