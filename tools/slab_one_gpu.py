"""Two slabs of the weak-scaling channel on ONE GPU (SlabGroup, peer-to-peer mailboxes inside one process): a target for
ncu captures of the exchange kernels (k_slab_edge_send, k_slab_unpack) on a single-GPU box.

    ncu --set full -k regex:k_slab -s 8 -c 4 -o prof python tools/slab_one_gpu.py [--steps 6]
"""
import argparse
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--DL", type=float, default=3.0, help="channel length per slab")
    a = ap.parse_args()
    from conftest import load_product
    load_product()
    cs = importlib.import_module("sph_b200.channel_setup")
    slab = importlib.import_module("sph_b200.slab")
    cfg = cs.ChannelConfig(DL=2 * a.DL, dp=0.0025, sort_interval=0).derive()
    pt = cs.cell_sorted(cfg, cs.developed(cfg, seed=12345))
    grp = slab.SlabGroup(cfg, pt, 2, compact_every=0, p2p=True)
    for _ in range(a.steps):
        grp.step()
    st = grp.status()
    print("steps", st["step"], "t", st["t"], "flags", st["error_flags"])
    grp.close()


if __name__ == "__main__":
    main()
