"""x-slab sharding of one long periodic channel over several GPUs: host side.

The reference has no domain decomposition.  The periodic channel is cut on cell-column
boundaries (cell = 2h, sph_neighbor_search_mex.c:253); each rank runs the device-resident step of
libsph_b200 on its slab plus four halo columns, and once per step -- before the list rebuild -- moves
the rows that changed slab and refreshes the halo copies.  Everything that touches particle data is a
CUDA kernel of the library (sphb_ctx_slab_pack / _unpack, csrc/slab.cu); this module only
  * cuts the initial condition into slabs (numpy bookkeeping, like channel_setup.py),
  * owns the exchange buffers (torch CUDA tensors: device memory only) and
  * moves them between ranks: NCCL send/recv + all-reduce through torch.distributed (`SlabChannel`),
    or plain device copies between slabs living on ONE GPU (`SlabGroup`, used by the parity tests and
    wherever there are fewer GPUs than slabs).
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import Channel, ChannelParams, SphError, _check, _p, lib   # noqa: F401

HALO_COLS = 4        # list -> density -> KGC -> forces -> continuity: one column of validity each
REC = 16             # doubles per record
HDR = 16             # doubles of header
_IP = C.POINTER(C.c_int)


def n_columns(cfg) -> int:
    return max(1, int(math.ceil(cfg.DL / (2.0 * cfg.h))))     # nb:254


def partition_columns(ncx: int, world: int):
    """Balanced cut of the cell columns [0, ncx) into `world` contiguous slabs."""
    base, extra = divmod(ncx, world)
    cuts, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < extra else 0)
        cuts.append((lo, hi))
        lo = hi
    return cuts


def _columns(x, cfg, ncx):
    cell = 2.0 * cfg.h
    xw = x - np.floor(x / cfg.DL) * cfg.DL                     # nb:270
    c = np.floor(xw / cell).astype(np.int64)
    return np.clip(c, 0, ncx - 1)


def cut_slab(cfg, pt, col_lo: int, col_hi: int, halo: int = HALO_COLS):
    """Owned fluid rows of the slab and the (static) wall rows of the slab and its halo columns, the
    latter unwrapped around the slab (x -+ DL across the periodic seam, as nb:284,291)."""
    ncx = n_columns(cfg)
    nf = pt.n_fluid
    c = _columns(pt.pos[:, 0], cfg, ncx)
    own = np.nonzero((c[:nf] >= col_lo) & (c[:nf] < col_hi))[0]
    width = col_hi - col_lo
    cw = c[nf:]
    rel = np.mod(cw - col_lo + halo, ncx) - halo
    keep = rel < width + halo
    if col_lo == 0:
        # towards the periodic seam the halo is `halo` CELLS of x, not columns: the channel's last column is a
        # sliver unless DL is a whole number of cells (same rule as slab_class in csrc/slab.cu for the fluid copies)
        xw = pt.pos[nf:, 0] - np.floor(pt.pos[nf:, 0] / cfg.DL) * cfg.DL
        deeper = ~keep & (xw >= cfg.DL - halo * 2.0 * cfg.h)
        rel = np.where(deeper, cw - ncx, rel)
        keep = keep | deeper
    k = (col_lo + rel - cw) // ncx                             # -1, 0 or +1 periods
    wx = pt.pos[nf:, 0] + k * cfg.DL
    widx = np.nonzero(keep)[0]
    wall_pos = np.asfortranarray(np.stack([wx[widx], pt.pos[nf:, 1][widx]], axis=1))
    return dict(
        ids=own.astype(np.int32),
        pos=np.asfortranarray(pt.pos[own, :]), vel=np.asfortranarray(pt.vel[own, :]),
        rho=np.ascontiguousarray(pt.rho[own]), drho_dt=np.ascontiguousarray(pt.drho_dt[own]),
        mass=np.ascontiguousarray(pt.mass[own]),
        wall_ids=(widx + nf).astype(np.int32), wall_pos=wall_pos,
        wall_mass=np.ascontiguousarray(pt.mass[nf:][widx]),
        wall_wall_vel=np.asfortranarray(pt.wall_vel[nf:, :][widx, :]))


class Slab:
    """One slab: a device-resident context in slab mode plus its four exchange buffers."""

    def __init__(self, cfg, pt, col_lo, col_hi, device, halo=HALO_COLS, headroom=1.25, **overrides):
        import torch
        self.torch = torch
        self.cfg, self.col_lo, self.col_hi, self.halo = cfg, col_lo, col_hi, halo
        self.device = torch.device(device)
        part = cut_slab(cfg, pt, col_lo, col_hi, halo)
        n_owned, n_wall = part["pos"].shape[0], part["wall_pos"].shape[0]
        width = col_hi - col_lo
        per_col = max(1.0, pt.n_fluid / n_columns(cfg))
        self.nf_cap = int((n_owned + 2 * (halo + 1) * per_col) * headroom) + 1024
        self.rec_cap = int((halo + 2) * per_col * 1.5) + 256
        kw = dict(n_fluid=self.nf_cap, n_total=self.nf_cap + n_wall, DL=cfg.DL, DH=cfg.DH, dp=cfg.dp, h=cfg.h,
                  rho0=cfg.rho0, p0=cfg.p0, c_f=cfg.c_f, mu=cfg.mu, gravity_g=cfg.gravity_g,
                  inv_sigma0=cfg.inv_sigma0, transport_coeff=cfg.transport_coeff, sort_interval=0,
                  device=self.device.index if self.device.index is not None else -1)
        kw.update(overrides)
        self.ch = Channel(ChannelParams(**kw))
        self._keep = part
        ids = part["ids"]
        _check(lib().sphb_ctx_slab_init(
            self.ch._ctx, C.c_int(col_lo), C.c_int(col_hi), C.c_int(halo), C.c_int(n_owned), _p(part["pos"]),
            _p(part["vel"]), _p(part["rho"]), _p(part["drho_dt"]), _p(part["mass"]), ids.ctypes.data_as(_IP),
            C.c_int(n_wall), _p(part["wall_pos"]), _p(part["wall_mass"]), _p(part["wall_wall_vel"])))
        n = HDR + self.rec_cap * REC
        mk = lambda: torch.zeros(n, dtype=torch.float64, device=self.device)   # noqa: E731
        self.send_left, self.send_right, self.recv_left, self.recv_right = mk(), mk(), mk(), mk()
        self.scal = torch.zeros(4, dtype=torch.float64, device=self.device)
        self.width = width

    # -- the two halves of a step and the device side of the exchange ---------------------------------
    def step_begin(self, remain=float("inf")):
        _check(lib().sphb_ctx_slab_step_begin(self.ch._ctx, C.c_double(remain)))

    def step_end(self):
        _check(lib().sphb_ctx_slab_step_end(self.ch._ctx))

    def pack(self):
        L = lib()
        _check(L.sphb_ctx_slab_scalars_export(self.ch._ctx, C.c_void_p(self.scal.data_ptr())))
        _check(L.sphb_ctx_slab_pack(self.ch._ctx, C.c_void_p(self.send_left.data_ptr()),
                                    C.c_void_p(self.send_right.data_ptr()), C.c_int(self.rec_cap)))

    def compact(self):
        _check(lib().sphb_ctx_slab_compact(self.ch._ctx))

    def unpack(self, import_vmax=True):
        L = lib()
        _check(L.sphb_ctx_slab_unpack(self.ch._ctx, C.c_void_p(self.recv_left.data_ptr()),
                                      C.c_void_p(self.recv_right.data_ptr()), C.c_void_p(self.send_left.data_ptr()),
                                      C.c_void_p(self.send_right.data_ptr()), C.c_int(self.rec_cap)))
        if import_vmax:
            _check(L.sphb_ctx_slab_vmax_import(self.ch._ctx, C.c_void_p(self.scal.data_ptr())))

    # -- peer-to-peer exchange (mailboxes; csrc/slab.cu) ------------------------------------------------------------
    def p2p_create(self, rank: int, world: int):
        """Allocates this slab's mailbox; returns (device pointer, 64-byte CUDA IPC handle)."""
        ptr, nbytes = C.c_void_p(0), C.c_size_t(0)
        handle = (C.c_ubyte * 64)()
        _check(lib().sphb_ctx_slab_p2p_create(self.ch._ctx, C.c_int(rank), C.c_int(world), C.c_int(self.rec_cap),
                                              C.byref(ptr), C.byref(nbytes), handle))
        self.mailbox = ptr.value
        return ptr.value, bytes(handle)

    def p2p_connect(self, rank: int, handle: bytes = None, pointer: int = None, peer_device: int = -1):
        h = (C.c_ubyte * 64).from_buffer_copy(handle) if handle is not None else None
        _check(lib().sphb_ctx_slab_p2p_connect(self.ch._ctx, C.c_int(rank), h, C.c_void_p(pointer or 0),
                                               C.c_int(peer_device)))

    def p2p_send(self):
        _check(lib().sphb_ctx_slab_p2p_send(self.ch._ctx, C.c_void_p(self.send_left.data_ptr()),
                                            C.c_void_p(self.send_right.data_ptr())))

    def p2p_recv(self):
        _check(lib().sphb_ctx_slab_p2p_recv(self.ch._ctx, C.c_void_p(self.send_left.data_ptr()),
                                            C.c_void_p(self.send_right.data_ptr())))

    def p2p_monitors(self):
        out = (C.c_double * 3)()
        _check(lib().sphb_ctx_slab_p2p_monitors(self.ch._ctx, out))
        return tuple(out)

    def rows_io(self, to_host: bool, n_rows: int, host: dict):
        """pos, vel, rho, drho_dt of the first n_rows owned-region rows to / from the (pinned) arrays in `host`."""
        _check(lib().sphb_ctx_slab_rows_io(self.ch._ctx, C.c_int(1 if to_host else 0), C.c_int(n_rows), _p(host["pos"]),
                                           _p(host["vel"]), _p(host["rho"]), _p(host["drho_dt"])))

    def counts(self):
        """(owned region incl. rows that have left and await compaction, halo copies, wall rows)."""
        a, b, c = C.c_int(0), C.c_int(0), C.c_int(0)
        _check(lib().sphb_ctx_slab_counts(self.ch._ctx, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def owned(self, fields=("pos", "vel", "rho", "p", "drho_dt")):
        """Owned fluid rows with their global ids: {"ids": ..., field: ...}."""
        n_owned, n_halo, n_wall = self.counts()
        out = self.ch.download(fields)              # capacity-sized arrays (the context is sized for its capacity)
        ids = np.zeros(n_owned + n_halo, dtype=np.int32)
        _check(lib().sphb_ctx_slab_ids(self.ch._ctx, ids.ctypes.data_as(_IP)))
        alive = ids[:n_owned] >= 0                    # rows that migrated away stay in place, marked -1
        res = {k: v[:n_owned][alive] for k, v in out.items()}
        res["ids"] = ids[:n_owned][alive]
        return res

    def reload(self, rows):
        """Re-initialise the slab from host rows (as returned by owned()): the H2D leg of an end-to-end step."""
        n_owned = rows["pos"].shape[0]
        keep = dict(ids=np.ascontiguousarray(rows["ids"], dtype=np.int32),
                    pos=np.asfortranarray(rows["pos"]), vel=np.asfortranarray(rows["vel"]),
                    rho=np.ascontiguousarray(rows["rho"]), drho_dt=np.ascontiguousarray(rows["drho_dt"]))
        mass = np.full(n_owned, self.cfg.rho0 * self.cfg.dp * self.cfg.dp) if "mass" not in rows else rows["mass"]
        keep["mass"] = np.ascontiguousarray(mass)
        part = self._keep
        n_wall = part["wall_pos"].shape[0]
        self.ch.prm.n_fluid, self.ch.prm.n_total = self.nf_cap, self.nf_cap + n_wall
        self._rows = keep
        _check(lib().sphb_ctx_slab_init(
            self.ch._ctx, C.c_int(self.col_lo), C.c_int(self.col_hi), C.c_int(self.halo), C.c_int(n_owned),
            _p(keep["pos"]), _p(keep["vel"]), _p(keep["rho"]), _p(keep["drho_dt"]), _p(keep["mass"]),
            keep["ids"].ctypes.data_as(_IP), C.c_int(n_wall), _p(part["wall_pos"]), _p(part["wall_mass"]),
            _p(part["wall_wall_vel"])))

    def close(self):
        self.ch.close()


class SlabGroup:
    """All slabs of a channel on ONE GPU, exchanged by device copies: the single-process form of the
    sharded run (parity tests; fewer GPUs than slabs).  Same kernels, same buffers as the NCCL form."""

    def __init__(self, cfg, pt, n_slabs: int, device="cuda:0", compact_every: int = 64, p2p: bool = False, **kw):
        cuts = partition_columns(n_columns(cfg), n_slabs)
        self.slabs = [Slab(cfg, pt, lo, hi, device, **kw) for lo, hi in cuts]
        self.n = n_slabs
        self.compact_every, self._k = compact_every, 0
        self.tau = (0.0, 0.0)
        self.p2p = p2p
        if p2p:      # mailboxes of one process: plain device pointers, no IPC
            boxes = [s.p2p_create(r, n_slabs)[0] for r, s in enumerate(self.slabs)]
            for s in self.slabs:
                for r, ptr in enumerate(boxes):
                    s.p2p_connect(r, pointer=ptr)
        self.exchange()
        for s in self.slabs:
            s.step_end()

    def exchange(self):
        import torch
        S = self.slabs
        self._k += 1
        if self.p2p:
            # every slab's pack kernel stores into its neighbours' mailboxes; the unpack kernels find them filled
            squeeze = self.compact_every and self._k % self.compact_every == 0
            for s in S:
                s.p2p_send()
                if squeeze:
                    s.compact()
            for s in S:
                s.ch.sync()          # (one GPU, one stream per slab: all sends before any receive)
            for s in S:
                s.p2p_recv()
            self.tau = S[0].p2p_monitors()[1:3]
            return
        for s in S:
            s.pack()
            if self.compact_every and self._k % self.compact_every == 0:
                s.compact()                # squeeze out the rows that have left since the last compaction
        for s in S:
            s.ch.sync()
        scal = torch.stack([s.scal for s in S])
        vmax = scal[:, 0].max()
        tau = scal[:, 1:3].sum(dim=0)
        for r, s in enumerate(S):
            left, right = S[(r - 1) % self.n], S[(r + 1) % self.n]
            s.recv_left.copy_(left.send_right)     # what my left neighbour sent to ITS right
            s.recv_right.copy_(right.send_left)
            s.scal[0] = vmax
        torch.cuda.synchronize()
        for s in S:
            s.unpack()
        self.tau = tuple(float(v) for v in tau.tolist())

    def step(self, remain=float("inf")):
        for s in self.slabs:
            s.step_begin(remain)
        self.exchange()
        for s in self.slabs:
            s.step_end()

    def monitors(self):
        """(max|v|, tau_bottom, tau_top) of the whole channel NOW: every slab's scalars, reduced on the device by the
        kernel the NCCL form uses on its gathered chunks (sphb_ctx_slab_scalars_reduce_at).  `self.tau` is what the
        last exchange carried, i.e. the shear of the list the step started from."""
        import torch
        S = self.slabs
        for s in S:
            _check(lib().sphb_ctx_slab_scalars_export(s.ch._ctx, C.c_void_p(s.scal.data_ptr())))
            s.ch.sync()
        gathered = torch.cat([s.scal for s in S])
        offs = torch.arange(0, 4 * self.n, 4, dtype=torch.int64, device=gathered.device)
        red = torch.zeros(3, dtype=torch.float64, device=gathered.device)
        _check(lib().sphb_ctx_slab_scalars_reduce_at(S[0].ch._ctx, C.c_void_p(gathered.data_ptr()),
                                                     C.c_void_p(offs.data_ptr()), C.c_int(self.n),
                                                     C.c_void_p(red.data_ptr())))
        S[0].ch.sync()
        return tuple(float(v) for v in red.tolist())   # (the kernel also stores max|v| in slab 0, which holds it already)

    def gather(self, fields=("pos", "vel", "rho", "p", "drho_dt")):
        """Owned rows of every slab, ordered by global particle id."""
        parts = [s.owned(fields) for s in self.slabs]
        ids = np.concatenate([p["ids"] for p in parts])
        order = np.argsort(ids, kind="stable")
        out = {k: np.concatenate([p[k] for p in parts], axis=0)[order] for k in fields}
        out["ids"] = ids[order]
        return out

    def status(self):
        st = [s.ch.status() for s in self.slabs]
        return dict(t=st[0]["t"], dt=st[0]["dt"], step=st[0]["step"], vmax=max(x["vmax"] for x in st),
                    error_flags=max(x["error_flags"] for x in st))

    def close(self):
        for s in self.slabs:
            s.close()


class NativeSlabGroup:
    """The sharded channel driven by the library's own C host (sphb_slabs_*, csrc/slab_group.cu): all slabs of a node in
    this process, one per GPU (or several per GPU), peer-to-peer mailboxes, one CUDA graph launch per GPU and step.
    No torch, no NCCL: what `sph_poiseuille_b200 --slabs N` runs.  Same interface as SlabGroup where it matters."""

    def __init__(self, cfg, pt, n_slabs: int, devices=None, compact_every: int = 256, **overrides):
        from . import _Params
        kw = dict(n_fluid=pt.n_fluid, n_total=pt.n_total, DL=cfg.DL, DH=cfg.DH, dp=cfg.dp, h=cfg.h, rho0=cfg.rho0,
                  p0=cfg.p0, c_f=cfg.c_f, mu=cfg.mu, gravity_g=cfg.gravity_g, inv_sigma0=cfg.inv_sigma0,
                  transport_coeff=cfg.transport_coeff, sort_interval=0, device=-1, max_neighbors_hint=0)
        kw.update(overrides)
        self.n_fluid, self.n_total = pt.n_fluid, pt.n_total
        cp = _Params(*[kw[k] for k in ("n_fluid", "n_total", "DL", "DH", "dp", "h", "rho0", "p0", "c_f", "mu", "gravity_g",
                                       "inv_sigma0", "transport_coeff", "sort_interval", "device", "max_neighbors_hint")])
        devs = None
        if devices is not None:
            devs = (C.c_int * len(devices))(*devices)
        f = np.asfortranarray
        self._g = C.c_void_p()
        _check(lib().sphb_slabs_create(C.byref(cp), C.c_int(n_slabs), devs, C.c_int(len(devices) if devices else 0),
                                       _p(f(pt.pos)), _p(f(pt.vel)), _p(np.ascontiguousarray(pt.rho)),
                                       _p(np.ascontiguousarray(pt.drho_dt)), _p(np.ascontiguousarray(pt.mass)),
                                       _p(f(pt.wall_vel)), C.byref(self._g)))
        _check(lib().sphb_slabs_set_compaction(self._g, C.c_int(compact_every)))
        self.n = n_slabs

    def step(self, remain=float("inf")):
        _check(lib().sphb_slabs_step(self._g, C.c_double(remain)))

    def run(self, n_steps: int):
        _check(lib().sphb_slabs_run(self._g, C.c_longlong(n_steps)))

    def run_until(self, t_target: float, t_end: float, max_steps: int = 0) -> int:
        n = C.c_longlong(0)
        _check(lib().sphb_slabs_run_until(self._g, C.c_double(t_target), C.c_double(t_end), C.c_longlong(max_steps),
                                          C.byref(n)))
        return n.value

    def sync(self):
        _check(lib().sphb_slabs_sync(self._g))

    def status(self):
        from . import _Status
        st = _Status()
        _check(lib().sphb_slabs_status(self._g, C.byref(st)))
        return {k: getattr(st, k) for k, _ in _Status._fields_}

    def gather(self, fields=("pos", "vel", "rho", "p", "drho_dt")):
        """Fluid rows by global particle id (the order of the arrays given at creation)."""
        nt = self.n_total
        buf = {k: (np.zeros((nt, 2), order="F") if k in ("pos", "vel") else np.zeros(nt)) for k in fields}
        arg = [_p(buf[k]) if k in buf else None for k in ("pos", "vel", "rho", "p", "drho_dt")]
        _check(lib().sphb_slabs_download(self._g, *arg))
        out = {k: v[:self.n_fluid] for k, v in buf.items()}
        out["ids"] = np.arange(self.n_fluid)
        return out

    def member_device(self, r: int) -> int:
        lo, hi, dev = C.c_int(0), C.c_int(0), C.c_int(0)
        L = lib()
        L.sphb_slabs_member.restype = C.c_void_p
        L.sphb_slabs_member(self._g, C.c_int(r), C.byref(lo), C.byref(hi), C.byref(dev))
        return dev.value

    def close(self):
        if self._g:
            L = lib()
            L.sphb_slabs_destroy.restype = None
            L.sphb_slabs_destroy(self._g)
            self._g = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def exchange_layout(rank: int, world: int, nbuf: int, n_scal: int = 4):
    """Layout of the one uneven all-to-all of a sharded step.  The chunk rank `rank` sends to / receives from
    rank d is [n_scal scalars][its "to the left" record buffer if d is its left neighbour][its "to the right"
    buffer if d is its right neighbour] (symmetric sizes).  Returns the split sizes, the chunk offsets, and where
    -- (offset, length) in the flat send / receive tensors -- the four record buffers live.  With two ranks
    the left and the right neighbour are the same peer and both buffers travel in its chunk: the peer's "to
    its left" buffer comes first and is what this rank receives FROM ITS RIGHT."""
    left, right = (rank - 1) % world, (rank + 1) % world
    split = [n_scal + (nbuf if d == left else 0) + (nbuf if d == right else 0) for d in range(world)]
    offs = [0]
    for n in split:
        offs.append(offs[-1] + n)

    def part(d, which):
        return offs[d] + n_scal + which * nbuf, nbuf
    if world == 1:
        lay = dict(send_left=None, send_right=None, recv_left=None, recv_right=None)
    elif left == right:
        lay = dict(send_left=part(left, 0), send_right=part(left, 1), recv_right=part(left, 0), recv_left=part(left, 1))
    else:
        lay = dict(send_left=part(left, 0), send_right=part(right, 0), recv_left=part(left, 0), recv_right=part(right, 0))
    lay.update(split=split, offs=offs[:-1], total=offs[-1], left=left, right=right)
    return lay


class SlabChannel:
    """One rank of the sharded channel: its slab, and the exchange over torch.distributed (backend "nccl").
    The two record buffers and the rank's scalars travel in ONE uneven all-to-all per step (NCCL send/recv
    over NVLink underneath): the left / right neighbour gets a record buffer + the scalars, every other rank
    the scalars only.  The library launches on the stream torch uses, so its kernels and the collective are
    stream-ordered, and the row counts stay on the device: a step has no host synchronisation and can be
    replayed as one CUDA graph (use_graph)."""

    def __init__(self, cfg, pt, rank: int, world: int, device, compact_every: int = 256, use_graph: bool = False,
                 p2p: bool = False, **kw):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank, self.world = rank, world
        self.compact_every, self._k = compact_every, 0
        self.use_graph, self._graph, self._graph_ok = use_graph, None, True
        lo, hi = partition_columns(n_columns(cfg), world)[rank]
        self.slab = Slab(cfg, pt, lo, hi, device, **kw)
        s = self.slab
        self.left, self.right = (rank - 1) % world, (rank + 1) % world
        self.stream = torch.cuda.Stream(device=s.device)     # shared by the library's kernels and torch / NCCL
        _check(lib().sphb_ctx_set_stream(s.ch._ctx, C.c_void_p(self.stream.cuda_stream)))
        self._red = torch.zeros(3, dtype=torch.float64, device=s.device)   # max|v|, sum tau_bottom, sum tau_top
        lay = exchange_layout(rank, world, s.send_left.numel())
        self._split = lay["split"]
        offs = np.asarray(lay["offs"], dtype=np.int64)
        self._xin = torch.zeros(lay["total"], dtype=torch.float64, device=s.device)
        self._xout = torch.zeros(lay["total"], dtype=torch.float64, device=s.device)
        self._scal_off = torch.tensor(offs, dtype=torch.int64, device=s.device)
        self._head_idx = torch.tensor(np.concatenate([o + np.arange(3) for o in offs]), dtype=torch.int64,
                                      device=s.device)
        self.p2p = bool(p2p) and world > 1
        if self.p2p:
            # mailboxes: every rank's receive area, mapped into every other rank through CUDA IPC; from then on a step
            # needs neither NCCL nor the host (the kernels of the library talk over NVLink)
            _, handle = s.p2p_create(rank, world)
            handles = [None] * world
            dist.all_gather_object(handles, handle)
            for r, h in enumerate(handles):
                if r != rank:
                    s.p2p_connect(r, handle=h)
            dist.barrier()
        elif world > 1:
            view = lambda t, ol: t[ol[0]:ol[0] + ol[1]]   # noqa: E731
            s.send_left, s.send_right = view(self._xin, lay["send_left"]), view(self._xin, lay["send_right"])
            s.recv_left, s.recv_right = view(self._xout, lay["recv_left"]), view(self._xout, lay["recv_right"])
        self.exchange()
        s.step_end()

    @property
    def tau(self):
        if self.p2p:
            return self.slab.p2p_monitors()[1:3]
        return tuple(float(v) for v in self._red[1:3].tolist())

    def exchange(self):
        with self.torch.cuda.stream(self.stream):
            self._exchange()

    def _exchange(self):
        dist, s = self.dist, self.slab
        self._k += 1
        squeeze = self.compact_every and self._k % self.compact_every == 0
        if self.p2p:
            s.p2p_send()
            if squeeze:
                s.compact()
            s.p2p_recv()
            return
        s.pack()
        if self.world == 1:
            if squeeze:
                s.compact()
            s.recv_left.copy_(s.send_right)
            s.recv_right.copy_(s.send_left)
            s.unpack(import_vmax=True)
            return
        self._xin.index_copy_(0, self._head_idx, s.scal[0:3].repeat(self.world))     # my scalars head every chunk
        work = dist.all_to_all_single(self._xout, self._xin, self._split, self._split, async_op=True)
        if squeeze:
            s.compact()                   # now and then: squeeze out the rows that have left (overlaps the transfer)
        work.wait()
        # max|v| -> time step of the next step (m:519-527); wall-shear partials (ph:1713-1742)
        _check(lib().sphb_ctx_slab_scalars_reduce_at(s.ch._ctx, C.c_void_p(self._xout.data_ptr()),
                                                     C.c_void_p(self._scal_off.data_ptr()), C.c_int(self.world),
                                                     C.c_void_p(self._red.data_ptr())))
        s.unpack(import_vmax=False)

    def _body(self, remain):
        self.slab.step_begin(remain)
        self._exchange()
        self.slab.step_end()

    def step(self, remain=float("inf")):
        """One reference step of the sharded channel.  With use_graph the whole step -- the library's kernels and
        the NCCL all-to-all, all on one stream, row counts on the device -- is captured once as a CUDA graph
        and replayed (remain = inf only); the now-and-then compaction step runs eagerly and the graph is
        captured again afterwards (the compaction swaps the state buffers)."""
        torch = self.torch
        squeeze_next = self.compact_every and (self._k + 1) % self.compact_every == 0
        if squeeze_next:
            # the compaction step runs eagerly anyway: look at the device error flags now (a slab whose list or
            # exchange buffers overflowed has been standing still since -- the kernels of a step return at once while
            # a flag is up -- and must not be mistaken for a running one)
            st = self.slab.ch.status()
            if st["error_flags"]:
                raise SphError("SPHB:capacity", f"slab {self.rank}: device error flags {st['error_flags']} at step "
                                                f"{st['step']} (list, tile or exchange-buffer capacity); the slab stands "
                                                f"at its last consistent step")
        if not (self.use_graph and self._graph_ok and remain == float("inf") and self.world > 1) or squeeze_next:
            with torch.cuda.stream(self.stream):
                self._body(remain)
            if squeeze_next:
                self._graph = None
            return
        if self._graph is None:
            k0 = self._k
            try:
                self.stream.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=self.stream):
                    self._body(remain)
                self._graph = g
                self._k = k0 + 1          # the capture enqueued nothing: replay it as this step
            except Exception:             # capture not possible here: stay eager
                self._graph_ok, self._graph, self._k = False, None, k0
                with torch.cuda.stream(self.stream):
                    self._body(remain)
                return
        else:
            self._k += 1
        with torch.cuda.stream(self.stream):
            self._graph.replay()

    def close(self):
        """Drop the captured graph (it holds NCCL work: do this BEFORE destroying the process group) and the slab."""
        self._graph = None
        self.slab.close()

    def __del__(self):
        self._graph = None
