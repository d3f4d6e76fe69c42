"""Sharded channel through the library's C host (sphb_slabs_*: ONE process, one slab per visible GPU, peer-to-peer
mailboxes, one CUDA graph launch per GPU and step) -- the same weak-scaling workload as `bench.py --gpus N`, which runs
one process per GPU under torchrun.  Prints one JSON line in bench.py's format.

    python tools/bench_native_slabs.py --slabs 8 [--DL 3] [--steps 100] [--warmup 10]

Timing: host clock around a region that starts and ends with every device synchronised (the steps in between are
enqueued without any host wait); value = particles of the whole channel x steps / that time.
"""
import argparse
import importlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--slabs", type=int, default=2)
    ap.add_argument("--DL", type=float, default=3.0, help="channel length per slab")
    ap.add_argument("--dp", type=float, default=0.0025)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    a = ap.parse_args()
    from conftest import load_product
    load_product()
    cs = importlib.import_module("sph_b200.channel_setup")
    slab = importlib.import_module("sph_b200.slab")
    cfg = cs.ChannelConfig(DL=a.DL * a.slabs, dp=a.dp, sort_interval=0).derive()
    pt = cs.cell_sorted(cfg, cs.developed(cfg, seed=12345))
    grp = slab.NativeSlabGroup(cfg, pt, a.slabs, compact_every=256)
    devs = [grp.member_device(r) for r in range(a.slabs)]
    grp.run(a.warmup)
    grp.sync()
    t0 = time.perf_counter()
    grp.run(a.steps)
    grp.sync()
    dt = time.perf_counter() - t0
    st = grp.status()
    n = pt.n_total
    print(json.dumps({
        "metric": "particle-substeps/s", "value": n * a.steps / dt, "unit": "particle-substeps/s", "n_gpus": len(set(devs)),
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * dt / a.steps, "higher_is_better": True, "scaling": "weak",
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"dp={a.dp} DL={cfg.DL:g} DH=1 channel, {pt.n_fluid}+{pt.n_total - pt.n_fluid} particles, "
                               f"x-slab sharded over {a.slabs} slabs on GPUs {devs}",
                   "host": "sphb_slabs_run: one process, one host thread, one CUDA graph launch per GPU and step, "
                           "peer-to-peer mailboxes (no NCCL, no torch)",
                   "timing": "host clock around a device-synchronised region", "l2": "no flush between steps"},
        "final_state": {"t": st["t"], "step": st["step"], "vmax": st["vmax"], "error_flags": st["error_flags"]}}))
    grp.close()


if __name__ == "__main__":
    main()
