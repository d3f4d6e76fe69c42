"""Phase times of one sharded step (x-slab channel, one rank per GPU): where the time beyond the single-GPU step goes.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/slab_phases.py \
        [--DL 3] [--nccl-exchange] [--steps 40]

Eager launches with CUDA events between the phases on the library's stream (so launch gaps are inside the figures; the
bench replays the same sequence as one CUDA graph and is faster by those gaps):

    physics    sphb_ctx_slab_step_begin   publish dt, density, KGC, fused forces, continuity
    send       pack (count, offsets, scatter) [+ the stores into the neighbours' mailboxes, peer-to-peer form]
    transfer   NCCL all-to-all (only with --nccl-exchange)
    recv       unpack (wait for the mailboxes, append rows, walls, commit counts / scalars)
    list       sphb_ctx_slab_step_end     bin count, scan, scatter, walk, wall shear

Prints the mean over steps of every rank's phase times (rank 0 prints the table, max and min over ranks).
"""
import argparse
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--DL", type=float, default=3.0, help="channel length per rank")
    ap.add_argument("--dp", type=float, default=0.0025)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--nccl-exchange", action="store_true")
    a = ap.parse_args()
    from conftest import load_product
    load_product()
    cs = importlib.import_module("sph_b200.channel_setup")
    slab = importlib.import_module("sph_b200.slab")
    rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    cfg = cs.ChannelConfig(DL=a.DL * world, dp=a.dp, sort_interval=0).derive()
    pt = cs.cell_sorted(cfg, cs.developed(cfg, seed=12345))
    sc = slab.SlabChannel(cfg, pt, rank, world, torch.device("cuda", lr), p2p=not a.nccl_exchange, compact_every=0)
    for _ in range(5):
        sc.step()
    K = a.steps
    names = ["physics", "send", "transfer", "recv", "list"]
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(6)] for _ in range(K)]
    torch.cuda.synchronize()
    dist.barrier()
    s = sc.slab
    L, C = slab.lib(), slab.C
    for k in range(K):
        with torch.cuda.stream(sc.stream):
            e = ev[k]
            e[0].record(sc.stream)
            s.step_begin()
            e[1].record(sc.stream)
            if sc.p2p:
                s.p2p_send()
                e[2].record(sc.stream)
                e[3].record(sc.stream)
                s.p2p_recv()
            else:
                s.pack()
                sc._xin.index_copy_(0, sc._head_idx, s.scal[0:3].repeat(world))
                e[2].record(sc.stream)
                dist.all_to_all_single(sc._xout, sc._xin, sc._split, sc._split)
                e[3].record(sc.stream)
                slab._check(L.sphb_ctx_slab_scalars_reduce_at(s.ch._ctx, C.c_void_p(sc._xout.data_ptr()),
                                                               C.c_void_p(sc._scal_off.data_ptr()), C.c_int(world),
                                                               C.c_void_p(sc._red.data_ptr())))
                s.unpack(import_vmax=False)
            e[4].record(sc.stream)
            s.step_end()
            e[5].record(sc.stream)
    torch.cuda.synchronize()
    g = np.array([[ev[k][i].elapsed_time(ev[k][i + 1]) for i in range(5)] for k in range(K)])[5:].mean(0)
    st = s.ch.status()
    allg = [None] * world
    dist.all_gather_object(allg, (g.tolist(), st["error_flags"], s.counts()))
    if rank == 0:
        t = np.array([x[0] for x in allg])
        print(f"# {world} ranks, {'peer-to-peer mailboxes' if sc.p2p else 'NCCL all-to-all'}, DL per rank {a.DL}, dp {a.dp}; "
              f"rows per rank (owned, halo, wall) {allg[0][2]}; error flags {[x[1] for x in allg]}")
        print("phase      " + "".join(f"{n:>10s}" for n in names) + "       sum")
        for label, v in (("max (ms)", t.max(0)), ("min (ms)", t.min(0)), ("mean (ms)", t.mean(0))):
            print(f"{label:11s}" + "".join(f"{x:10.4f}" for x in v) + f"{v.sum():10.4f}")
    dist.barrier()
    sc.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
