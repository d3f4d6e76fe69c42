"""Host-side logic of the x-slab sharding (no GPU): column partition, slab cut, and the rank-to-rank
bookkeeping over torch.distributed with the gloo backend, world_size 2."""
import importlib
import os
import sys

import numpy as np
import pytest


@pytest.fixture(scope="module")
def mods(sphb):
    return importlib.import_module("sph_b200.channel_setup"), importlib.import_module("sph_b200.slab")


def test_partition_is_contiguous_and_balanced(mods):
    _, slab = mods
    for ncx in (7, 58, 116, 462, 14770):
        for world in (1, 2, 3, 4, 8):
            cuts = slab.partition_columns(ncx, world)
            assert cuts[0][0] == 0 and cuts[-1][1] == ncx
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            widths = [hi - lo for lo, hi in cuts]
            assert max(widths) - min(widths) <= 1


def test_cut_covers_every_particle_once_and_unwraps_the_wall_halo(mods):
    cs, slab = mods
    cfg = cs.ChannelConfig(DL=2.0, DH=1.0, dp=0.02, c_f=10.0).derive()
    pt = cs.developed(cfg, seed=3)
    ncx = slab.n_columns(cfg)
    cell = 2.0 * cfg.h
    seen = []
    for lo, hi in slab.partition_columns(ncx, 3):
        part = slab.cut_slab(cfg, pt, lo, hi)
        seen.append(part["ids"])
        c = np.floor(part["pos"][:, 0] / cell).astype(int)
        assert c.min() >= lo and c.max() < hi
        # wall rows: the slab and its halo columns, contiguous around the slab (x -+ DL across the seam)
        wc = np.floor(part["wall_pos"][:, 0] / cell).astype(int)
        assert wc.min() >= lo - slab.HALO_COLS and wc.max() < hi + slab.HALO_COLS
        gx = pt.pos[part["wall_ids"], 0]
        k = np.rint((part["wall_pos"][:, 0] - gx) / cfg.DL)
        assert set(np.unique(k)).issubset({-1.0, 0.0, 1.0})
        assert np.array_equal(part["wall_pos"][:, 0], gx + k * cfg.DL)       # the reference's ghost arithmetic
    ids = np.sort(np.concatenate(seen))
    assert np.array_equal(ids, np.arange(pt.n_fluid))


@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_all_to_all_layout_delivers_each_buffer_to_the_right_neighbour(mods, world):
    """The one uneven all-to-all of a sharded step, simulated: every rank's "to the left" buffer must arrive as
    its left neighbour's "from the right" and vice versa (two ranks: both buffers share one chunk)."""
    _, slab = mods
    nbuf = 5
    lays = [slab.exchange_layout(r, world, nbuf) for r in range(world)]
    xin = []
    for r, lay in enumerate(lays):
        x = np.full(lay["total"], -1.0)
        for d in range(world):
            x[lay["offs"][d]:lay["offs"][d] + 3] = 100 * r + np.arange(3)            # scalars head every chunk
        o, n = lay["send_left"]
        x[o:o + n] = 1000 * r + 1                                                    # tag: r's buffer for its left
        o, n = lay["send_right"]
        x[o:o + n] = 1000 * r + 2                                                    # tag: r's buffer for its right
        xin.append(x)
    for r, lay in enumerate(lays):
        # all_to_all_single: chunk d of the output of r = chunk r of the input of d
        out = np.concatenate([xin[d][lays[d]["offs"][r]:lays[d]["offs"][r] + lays[d]["split"][r]] for d in range(world)])
        assert out.size == lay["total"] and [lays[d]["split"][r] for d in range(world)] == lay["split"]
        o, n = lay["recv_left"]
        assert np.all(out[o:o + n] == 1000 * lay["left"] + 2)                        # my left neighbour's "to the right"
        o, n = lay["recv_right"]
        assert np.all(out[o:o + n] == 1000 * lay["right"] + 1)                       # my right neighbour's "to the left"
        for d in range(world):
            assert np.array_equal(out[lay["offs"][d]:lay["offs"][d] + 3], 100 * d + np.arange(3))


def _gloo_worker(rank, world, port, root, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    import torch
    import torch.distributed as dist
    from conftest import load_product
    load_product()
    cs = importlib.import_module("sph_b200.channel_setup")
    slab = importlib.import_module("sph_b200.slab")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = cs.ChannelConfig(DL=2.0, DH=1.0, dp=0.02, c_f=10.0).derive()
    pt = cs.developed(cfg, seed=3)
    ncx = slab.n_columns(cfg)
    lo, hi = slab.partition_columns(ncx, world)[rank]
    part = slab.cut_slab(cfg, pt, lo, hi)
    # every particle is owned exactly once over the ranks
    n = torch.tensor([part["ids"].size], dtype=torch.int64)
    dist.all_reduce(n)
    mark = torch.zeros(pt.n_fluid, dtype=torch.int64)
    mark[torch.from_numpy(part["ids"].astype(np.int64))] = 1
    dist.all_reduce(mark)
    # edge copies a rank would send to its right neighbour == what that neighbour expects on its left
    cell = 2.0 * cfg.h
    c = np.floor(part["pos"][:, 0] / cell).astype(int)
    to_right = np.sort(part["ids"][c >= hi - slab.HALO_COLS]).astype(np.int64)
    right, left = (rank + 1) % world, (rank - 1) % world
    n_send = torch.tensor([to_right.size], dtype=torch.int64)
    n_recv = torch.zeros(1, dtype=torch.int64)
    reqs = [dist.isend(n_send, right), dist.irecv(n_recv, left)]
    for r in reqs:
        r.wait()
    buf = torch.zeros(int(n_recv.item()), dtype=torch.int64)
    reqs = [dist.isend(torch.from_numpy(to_right), right), dist.irecv(buf, left)]
    for r in reqs:
        r.wait()
    # what arrived must be the global particles of the H columns left of my slab (periodic)
    cg = np.floor(pt.pos[:pt.n_fluid, 0] / cell).astype(int)
    want = np.nonzero(np.mod(cg - (lo - slab.HALO_COLS), ncx) < slab.HALO_COLS)[0]
    ok = int(n.item()) == pt.n_fluid and bool((mark == 1).all()) and np.array_equal(np.sort(buf.numpy()), want)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        open(out, "w").write("ok" if ok else "bad")
    if not ok:
        raise SystemExit(3)


def test_two_ranks_agree_on_ownership_and_halo_over_gloo(tmp_path):
    import torch.multiprocessing as mp
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = str(tmp_path / "result.txt")
    port = 29600 + os.getpid() % 300
    mp.spawn(_gloo_worker, args=(2, port, root, out), nprocs=2, join=True)
    assert open(out).read() == "ok"


def test_cell_sorted_matches_the_reference_sort(mods, H):
    """channel_setup.cell_sorted (the row order the benchmark starts from) is the reference's
    sort_particles_by_cell (SPH_Poiseuille.m:529-568) as restated by the oracle host."""
    cs, _ = mods
    cfg = cs.ChannelConfig(DL=3.0, DH=1.0, dp=0.05, c_f=10.0).derive()
    pt = cs.developed(cfg, seed=17, jitter=0.3)
    got = cs.cell_sorted(cfg, pt)
    perm = H.sort_permutation(pt.pos, pt.n_fluid, pt.n_total, cfg.DL, cfg.h)
    assert np.array_equal(got.pos, pt.pos[perm]) and np.array_equal(got.vel, pt.vel[perm])
    assert np.array_equal(got.rho, pt.rho[perm]) and np.array_equal(got.drho_dt, pt.drho_dt[perm])
    assert got.n_fluid == pt.n_fluid and np.array_equal(got.pos[pt.n_fluid:], pt.pos[pt.n_fluid:])
