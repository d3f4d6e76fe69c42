"""oracle/host.py -- TEST INFRASTRUCTURE, not product code.

The "fake MATLAB host": the thin slice of SPH_Poiseuille.m that sits around the two MEX files,
restated in numpy so the CPU oracles can be driven through whole time steps without MATLAB.

  * :func:`derive_params`      m:62-80   (geometry snapping, g, h, BW, p0, inv_sigma0, nu)
  * :func:`init_particles`     m:95-123 + build_shell_wall_particles.m:14-38
  * :func:`verlet_time_step`   m:519-527
  * :func:`periodic_bounding`  m:570-577
  * :func:`sort_permutation`   m:529-568
  * :func:`reference_step`     m:250-283 (one pass of the main loop body, same call order)
  * :func:`mid_channel_profile` / :func:`binned_profile` / :func:`l2_error`  m:579-605, postprocess:42

Two interchangeable back ends expose the MEX call surface (``neighbor_search`` / ``physics``):

  * :class:`RefBackend`  -- oracle/_ref/libsph_ref.so: the reference's C sources compiled unmodified;
  * :class:`PortBackend` -- oracle/libsph_oracle.so: this repo's plain-C restatement.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm may import this.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from dataclasses import dataclass, field

import numpy as np

from .mexhost import MexLibrary, MexError  # noqa: F401  (re-exported)

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libsph_ref.so")
PORT_SO = os.path.join(HERE, "libsph_oracle.so")


# ------------------------------------------------------------------------------------------------
# parameters and initial condition
# ------------------------------------------------------------------------------------------------
@dataclass
class Params:
    DL: float
    DH: float
    dp: float
    rho0: float
    mu: float
    U_bulk: float
    c_f: float
    t_end: float = 5.0
    output_interval: float = 0.2
    sort_interval: int = 100
    transport_coeff: float = 0.30  # hard-coded in the script, m:77
    # derived (m:71-80)
    gravity_g: float = field(init=False)
    U_max: float = field(init=False)
    h: float = field(init=False)
    wall_thickness: float = field(init=False)
    p0: float = field(init=False)
    inv_sigma0: float = field(init=False)
    nu: float = field(init=False)

    def __post_init__(self):
        # m:64-65: snap DL, DH to multiples of dp (MATLAB round = half away from zero)
        self.DL = _matlab_round(self.DL / self.dp) * self.dp
        self.DH = _matlab_round(self.DH / self.dp) * self.dp
        self.gravity_g = 12.0 * self.mu * self.U_bulk / (self.rho0 * self.DH ** 2)
        self.U_max = 1.5 * self.U_bulk
        self.h = 1.3 * self.dp
        cutoff_depth = math.ceil((2.0 * self.h) / self.dp) * self.dp
        self.wall_thickness = max(4.0 * self.dp, cutoff_depth)
        self.p0 = self.rho0 * self.c_f ** 2
        self.inv_sigma0 = self.dp ** 2
        self.nu = self.mu / self.rho0


def _matlab_round(v: float) -> float:
    return math.floor(abs(v) + 0.5) * (1.0 if v >= 0 else -1.0)


def derive_params(DL=3.0, DH=1.0, dp=0.05, rho0=1.0, mu=0.1, U_bulk=0.666667, c_f=15.0, **kw) -> Params:
    return Params(DL=DL, DH=DH, dp=dp, rho0=rho0, mu=mu, U_bulk=U_bulk, c_f=c_f, **kw)


def _colon(a: float, d: float, n: int) -> np.ndarray:
    """a : d : a+(n-1)d.  MATLAB's colon may differ from a + k*d in the last ulp for some k; the MEX
    path never sees that (MATLAB hands it ``pos``), and both the oracle host and the product's C
    driver use a + k*d, so they agree with each other."""
    return a + d * np.arange(n, dtype=np.float64)


@dataclass
class State:
    pos: np.ndarray        # [n_total, 2] Fortran order
    vel: np.ndarray
    rho: np.ndarray        # [n_total]
    p: np.ndarray
    drho_dt: np.ndarray
    force: np.ndarray      # [n_total, 2]
    force_prior: np.ndarray
    Vol: np.ndarray
    B: np.ndarray          # [n_total, 4]
    mass: np.ndarray
    wall_vel: np.ndarray
    n_fluid: int
    n_total: int
    t: float = 0.0
    step: int = 0
    outer: int = 0          # outer-step counter of the dual-criteria schedule
    tau_bottom: float = 0.0
    tau_top: float = 0.0
    pairs: tuple | None = None   # (pair_i, pair_j, dx, dy, r, W, dW)

    def copy(self) -> "State":
        kw = {}
        for k, v in self.__dict__.items():
            if isinstance(v, np.ndarray):
                kw[k] = v.copy(order="F")
            elif isinstance(v, tuple):
                kw[k] = tuple(a.copy() for a in v)
            else:
                kw[k] = v
        return State(**kw)


def init_particles(prm: Params) -> State:
    """Lattice initial condition.  Fluid: meshgrid(x,y)(:) => y runs fastest (m:95-98).  Walls: bottom
    block then top block, each y-fastest (build_shell_wall_particles.m:22-31)."""
    dp = prm.dp
    nx = int(_matlab_round(prm.DL / dp))
    ny = int(_matlab_round(prm.DH / dp))
    nl = int(_matlab_round(prm.wall_thickness / dp))
    xf = _colon(dp / 2, dp, nx)
    yf = _colon(dp / 2, dp, ny)
    yb = _colon(-prm.wall_thickness + dp / 2, dp, nl)
    yt = _colon(prm.DH + dp / 2, dp, nl)

    def block(xs, ys):
        X = np.repeat(xs, len(ys))
        Y = np.tile(ys, len(xs))
        return np.stack([X, Y], axis=1)

    pos = np.concatenate([block(xf, yf), block(xf, yb), block(xf, yt)], axis=0)
    n_fluid = nx * ny
    n_total = pos.shape[0]
    mass = np.full(n_total, prm.rho0 * dp ** 2)
    mass[n_fluid:] = prm.rho0 * (dp * dp)  # wall_measure .* wall_thickness_arr = dp*dp (m:117-118)
    rho = np.full(n_total, prm.rho0)
    B = np.zeros((n_total, 4), order="F")
    B[:, 0] = 1.0
    B[:, 3] = 1.0
    z2 = lambda: np.zeros((n_total, 2), order="F")  # noqa: E731
    return State(pos=np.asfortranarray(pos), vel=z2(), rho=rho, p=np.zeros(n_total), drho_dt=np.zeros(n_total),
                 force=z2(), force_prior=z2(), Vol=mass / rho, B=B, mass=mass, wall_vel=z2(),
                 n_fluid=n_fluid, n_total=n_total)


def developed_state(prm: Params, seed: int = 12345, jitter: float = 0.1, noise: float = 0.01) -> State:
    """Seeded synthetic 'developed' state (SURVEY.md 8(d), config 3 kernel-test variant): lattice +
    uniform +-jitter*dp displacement of the fluid rows, analytic u(y) + noise*U_max velocity noise,
    small random drho_dt.  Deterministic for a given seed (numpy PCG64)."""
    st = init_particles(prm)
    rng = np.random.Generator(np.random.PCG64(seed))
    nf = st.n_fluid
    st.pos[:nf, :] += (rng.random((nf, 2)) * 2.0 - 1.0) * jitter * prm.dp
    y = st.pos[:nf, 1]
    u = prm.gravity_g / (2.0 * prm.nu) * y * (prm.DH - y)
    st.vel[:nf, 0] = u + noise * prm.U_max * (rng.random(nf) * 2.0 - 1.0)
    st.vel[:nf, 1] = noise * prm.U_max * (rng.random(nf) * 2.0 - 1.0)
    st.drho_dt[:nf] = 0.05 * prm.rho0 * (rng.random(nf) * 2.0 - 1.0)
    st.pos[:nf, 0] = np.mod(st.pos[:nf, 0], prm.DL)
    return st


# ------------------------------------------------------------------------------------------------
# back ends
# ------------------------------------------------------------------------------------------------
class RefBackend:
    """The reference MEX sources, compiled unmodified (oracle/_ref/libsph_ref.so)."""

    kind = "reference"

    def __init__(self, path: str = REF_SO):
        self.mex = MexLibrary(path)

    def neighbor_search(self, pos, n_fluid, n_total, h, DL):
        out = self.mex.call("ref_neighbor_mexFunction", 7, pos, n_fluid, n_total, h, DL)
        return tuple(o.ravel(order="F") for o in out)

    def physics(self, mode: str, *args, nargout: int):
        out = self.mex.call("ref_physics_mexFunction", nargout, mode, *args)
        return [o if o.shape[1] > 1 else o.ravel(order="F") for o in out]


_DP = C.POINTER(C.c_double)


def _p(a: np.ndarray):
    return a.ctypes.data_as(_DP)


def _f(a, shape=None) -> np.ndarray:
    arr = np.asfortranarray(np.asarray(a, dtype=np.float64))
    if shape is not None and arr.shape != shape:
        arr = np.asfortranarray(arr.reshape(shape, order="F"))
    return arr


class _OrcPairs(C.Structure):
    _fields_ = [("n_pairs", C.c_size_t)] + [(k, _DP) for k in ("pair_i", "pair_j", "dx", "dy", "r", "W", "dW")]


class PortBackend:
    """This repo's plain-C restatement (oracle/libsph_oracle.so), same call surface as the MEX files."""

    kind = "port"

    def __init__(self, path: str = PORT_SO):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        self.lib.orc_verlet_time_step.restype = C.c_double

    # -- neighbour search ------------------------------------------------------------------------
    def neighbor_search(self, pos, n_fluid, n_total, h, DL):
        pos = _f(pos)
        out = _OrcPairs()
        rc = self.lib.orc_neighbor_search(_p(pos), C.c_int(int(n_fluid)), C.c_int(int(n_total)),
                                          C.c_double(h), C.c_double(DL), C.byref(out))
        if rc == -1:
            raise MexError("SPH:Neighbor:count", "Invalid n_fluid/n_total or inconsistent pos size.")
        if rc == -2:
            raise MexError("SPH:Neighbor:param", "h and DL must be positive.")
        n = out.n_pairs
        res = tuple(np.ctypeslib.as_array(getattr(out, k), shape=(n,)).copy() if n else np.zeros(0)
                    for k in ("pair_i", "pair_j", "dx", "dy", "r", "W", "dW"))
        self.lib.orc_pairs_free(C.byref(out))
        return res

    # -- physics modes (argument order = the MEX modes') ---------------------------
    def physics(self, mode: str, *a, nargout: int):
        L = self.lib
        d = C.c_double
        ci = C.c_int
        if mode == "density_correction":
            pi, pj, dx, dy, r, W, dW, mass, nf, nt, rho0, h, isg = a
            nf, nt = int(nf), int(nt)
            arrs = [_f(v) for v in (pi, pj, dx, dy, r, W, dW, mass)]
            rho, Vol, B = np.zeros(nt), np.zeros(nt), np.zeros((nt, 4), order="F")
            L.orc_density_correction(C.c_size_t(arrs[0].size), *[_p(v) for v in arrs], ci(nf), ci(nt), d(rho0),
                                     d(h), d(isg), _p(rho), _p(Vol), _p(B))
            return [rho, Vol, B]
        if mode == "viscous_force":
            pi, pj, dx, dy, r, dW, vel, Vol, B, mu, h, nf, nt, mass, wv = a[:15]
            nf, nt = int(nf), int(nt)
            arrs = [_f(v) for v in (pi, pj, dx, dy, r, dW, vel, Vol, B)]
            wv = _f(wv)
            force = np.zeros((nt, 2), order="F")
            L.orc_viscous_force(C.c_size_t(arrs[0].size), *[_p(v) for v in arrs], d(mu), d(h), ci(nf), ci(nt),
                                _p(wv), _p(force))
            return [force]
        if mode == "transport_correction":
            pi, pj, dx, dy, r, dW, Vol, B, pos, h, nf, nt = a[:12]
            coeff = a[12] if len(a) > 12 else 0.2
            nf, nt = int(nf), int(nt)
            arrs = [_f(v) for v in (pi, pj, dx, dy, r, dW, Vol, B, pos)]
            out = np.zeros((nt, 2), order="F")
            L.orc_transport_correction(C.c_size_t(arrs[0].size), *[_p(v) for v in arrs], d(h), ci(nf), ci(nt),
                                       d(coeff), _p(out))
            return [out]
        if mode in ("integration_1st", "integration_verlet"):
            pi, pj, dx, dy, r, dW, Vol, B, rho, mass, pos, vel, drho, fp, dt, nf, nt, rho0, p0, c_f, wv = a
            nf, nt = int(nf), int(nt)
            arrs = [_f(v) for v in (pi, pj, dx, dy, r, dW, Vol, B, rho, mass, pos, vel, drho, fp)]
            wv = _f(wv)
            rho_o, p_o, drho_o = np.zeros(nt), np.zeros(nt), np.zeros(nt)
            pos_o, vel_o, f_o = (np.zeros((nt, 2), order="F") for _ in range(3))
            if mode == "integration_1st":
                L.orc_integration_1st(C.c_size_t(arrs[0].size), *[_p(v) for v in arrs], d(dt), ci(nf), ci(nt),
                                      d(rho0), d(p0), d(c_f), _p(rho_o), _p(p_o), _p(pos_o), _p(f_o), _p(drho_o))
                return [rho_o, p_o, pos_o, f_o, drho_o]
            L.orc_integration_verlet(C.c_size_t(arrs[0].size), *[_p(v) for v in arrs], d(dt), ci(nf), ci(nt),
                                     d(rho0), d(p0), d(c_f), _p(wv), _p(rho_o), _p(p_o), _p(pos_o), _p(vel_o),
                                     _p(drho_o), _p(f_o))
            return [rho_o, p_o, pos_o, vel_o, drho_o, f_o]
        if mode == "integration_2nd":
            pi, pj, dx, dy, r, dW, Vol, rho, pos, vel, dt, nf, nt, wv = a
            nf, nt = int(nf), int(nt)
            arrs = [_f(v) for v in (pi, pj, dx, dy, r, dW, Vol, rho, pos, vel)]
            wv = _f(wv)
            pos_o, drho_o = np.zeros((nt, 2), order="F"), np.zeros(nt)
            L.orc_integration_2nd(C.c_size_t(arrs[0].size), *[_p(v) for v in arrs], d(dt), ci(nf), ci(nt), _p(wv),
                                  _p(pos_o), _p(drho_o))
            return [pos_o, drho_o, np.zeros((nt, 2), order="F")]
        if mode == "advance_shell_step":
            (pi, pj, dx, dy, r, W, dW, mass, pos, vel, wv, rho, drho, dt, nf, nt, rho0, p0, c_f, mu, h, isg,
             g) = a
            nf, nt = int(nf), int(nt)
            arrs = [_f(v) for v in (pi, pj, dx, dy, r, W, dW, mass, pos, vel, wv, drho)]
            rho_o, p_o, drho_o, Vol_o = (np.zeros(nt) for _ in range(4))
            pos_o, vel_o, f_o, fp_o = (np.zeros((nt, 2), order="F") for _ in range(4))
            B_o = np.zeros((nt, 4), order="F")
            L.orc_advance_shell_step(C.c_size_t(arrs[0].size), *[_p(v) for v in arrs], d(dt), ci(nf), ci(nt),
                                     d(rho0), d(p0), d(c_f), d(mu), d(h), d(isg), d(g), _p(rho_o), _p(p_o),
                                     _p(pos_o), _p(vel_o), _p(drho_o), _p(f_o), _p(fp_o), _p(Vol_o), _p(B_o))
            return [rho_o, p_o, pos_o, vel_o, drho_o, f_o, fp_o, Vol_o, B_o]
        if mode == "wall_shear_monitor":
            pi, pj, dx, dy, r, dW, pos, vel, wv, Vol, B, nf, DL, DH, mu, h = a
            arrs = [_f(v) for v in (pi, pj, dx, dy, r, dW, pos, vel, wv, Vol, B)]
            nt = arrs[9].size
            tb, tt = d(0.0), d(0.0)
            L.orc_wall_shear_monitor(C.c_size_t(arrs[0].size), *[_p(v) for v in arrs], ci(int(nf)), ci(nt), d(DL),
                                     d(DH), d(mu), d(h), C.byref(tb), C.byref(tt))
            return [np.array([[tb.value]]), np.array([[tt.value]])]
        raise MexError("SPH:Physics:mode", "Unsupported mode.")


def available_backends():
    out = []
    if os.path.exists(REF_SO):
        out.append(RefBackend())
    if os.path.exists(PORT_SO):
        out.append(PortBackend())
    return out


def best_backend():
    """The reference build when present (oracle/_ref), else the port."""
    if os.path.exists(REF_SO):
        return RefBackend()
    return PortBackend()


# ------------------------------------------------------------------------------------------------
# the main-loop slice of SPH_Poiseuille.m
# ------------------------------------------------------------------------------------------------
def verlet_time_step(vel_fluid, c_max, h, nu, gravity_g, remain):  # m:519-527
    v_max = float(np.max(np.sqrt(vel_fluid[:, 0] ** 2 + vel_fluid[:, 1] ** 2)))
    dt_acoustic = 0.25 * h / max(c_max + v_max, 1e-12)
    dt_viscous = 0.125 * h * h / max(nu, 1e-12)
    dt_body = 0.25 * math.sqrt(h / max(abs(gravity_g), 1e-12))
    return max(min(dt_acoustic, dt_viscous, dt_body, remain), 1e-12)


def periodic_bounding(pos, n_fluid, DL):  # m:570-577
    pos[:n_fluid, 0] = np.mod(pos[:n_fluid, 0], DL)
    return pos


def sort_permutation(pos, n_fluid, n_total, DL, h):  # m:529-568
    if n_fluid <= 0 or n_fluid >= n_total:
        return np.arange(n_total)
    cell = 2.0 * h
    cx = np.floor(np.mod(pos[:n_fluid, 0], DL) / cell)
    cy = np.floor((pos[:n_fluid, 1] - pos[:n_fluid, 1].min()) / cell)
    idx_f = np.lexsort((cx, cy))  # stable, primary key cy, secondary cx (sortrows([cy,cx],[1,2]))
    return np.concatenate([idx_f, np.arange(n_fluid, n_total)])


def build_neighbor_cache(be, st: State, prm: Params):
    st.pairs = be.neighbor_search(st.pos, st.n_fluid, st.n_total, prm.h, prm.DL)
    return st.pairs


def initialise(be, st: State, prm: Params):
    """S5 of the script (m:165-173): first neighbour list, density/KGC, EOS."""
    pi, pj, dx, dy, r, W, dW = build_neighbor_cache(be, st, prm)
    st.rho, st.Vol, st.B = be.physics("density_correction", pi, pj, dx, dy, r, W, dW, st.mass, st.n_fluid,
                                      st.n_total, prm.rho0, prm.h, prm.inv_sigma0, nargout=3)
    st.p = np.zeros(st.n_total)
    st.p[:st.n_fluid] = prm.p0 * (st.rho[:st.n_fluid] / prm.rho0 - 1.0)
    return st


def reference_step(be, st: State, prm: Params, remain: float = math.inf, do_sort: bool = True,
                   transport_coeff: float | None = None):
    """One pass of the main-loop body, m:251-283, in the reference's call order.  Returns dt."""
    nf, nt = st.n_fluid, st.n_total
    coeff = prm.transport_coeff if transport_coeff is None else transport_coeff
    st.step += 1
    pi, pj, dx, dy, r, W, dW = st.pairs
    st.rho, st.Vol, st.B = be.physics("density_correction", pi, pj, dx, dy, r, W, dW, st.mass, nf, nt,
                                      prm.rho0, prm.h, prm.inv_sigma0, nargout=3)
    (fp,) = be.physics("viscous_force", pi, pj, dx, dy, r, dW, st.vel, st.Vol, st.B, prm.mu, prm.h, nf, nt,
                       st.mass, st.wall_vel, nargout=1)
    fp[:nf, 0] = fp[:nf, 0] + st.mass[:nf] * prm.gravity_g  # m:392
    fp[nf:, :] = 0.0
    st.force_prior = fp
    (st.pos,) = be.physics("transport_correction", pi, pj, dx, dy, r, dW, st.Vol, st.B, st.pos, prm.h, nf, nt,
                           coeff, nargout=1)
    dt = verlet_time_step(st.vel[:nf], prm.c_f, prm.h, prm.nu, prm.gravity_g, remain)
    st.rho, st.p, st.pos, st.vel, st.drho_dt, st.force = be.physics(
        "integration_verlet", pi, pj, dx, dy, r, dW, st.Vol, st.B, st.rho, st.mass, st.pos, st.vel, st.drho_dt,
        st.force_prior, dt, nf, nt, prm.rho0, prm.p0, prm.c_f, st.wall_vel, nargout=6)
    st.t += dt
    periodic_bounding(st.pos, nf, prm.DL)
    st.vel[nf:, :] = 0.0
    if do_sort and st.step % prm.sort_interval == 0 and st.step != 1:  # m:272
        perm = sort_permutation(st.pos, nf, nt, prm.DL, prm.h)
        for name in ("pos", "vel", "rho", "mass", "wall_vel", "drho_dt", "force_prior", "force", "p", "Vol", "B"):
            setattr(st, name, np.asfortranarray(getattr(st, name)[perm]))
    pi, pj, dx, dy, r, W, dW = build_neighbor_cache(be, st, prm)
    tb, tt = be.physics("wall_shear_monitor", pi, pj, dx, dy, r, dW, st.pos, st.vel, st.wall_vel, st.Vol, st.B,
                        nf, prm.DL, prm.DH, prm.mu, prm.h, nargout=2)
    st.tau_bottom, st.tau_top = float(np.ravel(tb)[0]), float(np.ravel(tt)[0])
    return dt


def dual_outer_step(be, st: State, prm: Params, n_inner: int, do_sort: bool = True):
    """One OUTER step of the dual-criteria schedule the reference README describes (README.md:61-65; no
    code at HEAD, so this is a composition of the pinned operators, not a pinned schedule): density +
    KGC, viscous (+ m g), transport on the current list, then `n_inner` integration_verlet substeps on
    the SAME (frozen) pair list, each with dt = verlet_time_step(...) and an immediate periodic wrap,
    then [cell sort], list rebuild, shear monitor.  n_inner = 1 is exactly reference_step(remain=inf).
    Returns the list of inner dts."""
    nf, nt = st.n_fluid, st.n_total
    pi, pj, dx, dy, r, W, dW = st.pairs
    st.rho, st.Vol, st.B = be.physics("density_correction", pi, pj, dx, dy, r, W, dW, st.mass, nf, nt,
                                      prm.rho0, prm.h, prm.inv_sigma0, nargout=3)
    (fp,) = be.physics("viscous_force", pi, pj, dx, dy, r, dW, st.vel, st.Vol, st.B, prm.mu, prm.h, nf, nt,
                       st.mass, st.wall_vel, nargout=1)
    fp[:nf, 0] = fp[:nf, 0] + st.mass[:nf] * prm.gravity_g
    fp[nf:, :] = 0.0
    st.force_prior = fp
    (st.pos,) = be.physics("transport_correction", pi, pj, dx, dy, r, dW, st.Vol, st.B, st.pos, prm.h, nf, nt,
                           prm.transport_coeff, nargout=1)
    dts = []
    for _ in range(n_inner):
        dt = verlet_time_step(st.vel[:nf], prm.c_f, prm.h, prm.nu, prm.gravity_g, math.inf)
        st.rho, st.p, st.pos, st.vel, st.drho_dt, st.force = be.physics(
            "integration_verlet", pi, pj, dx, dy, r, dW, st.Vol, st.B, st.rho, st.mass, st.pos, st.vel, st.drho_dt,
            st.force_prior, dt, nf, nt, prm.rho0, prm.p0, prm.c_f, st.wall_vel, nargout=6)
        st.t += dt
        st.step += 1
        periodic_bounding(st.pos, nf, prm.DL)
        st.vel[nf:, :] = 0.0
        dts.append(dt)
    st.outer += 1
    if do_sort and prm.sort_interval > 0 and st.outer % prm.sort_interval == 0 and st.outer != 1:
        perm = sort_permutation(st.pos, nf, nt, prm.DL, prm.h)
        for name in ("pos", "vel", "rho", "mass", "wall_vel", "drho_dt", "force_prior", "force", "p", "Vol", "B"):
            setattr(st, name, np.asfortranarray(getattr(st, name)[perm]))
    pi, pj, dx, dy, r, W, dW = build_neighbor_cache(be, st, prm)
    tb, tt = be.physics("wall_shear_monitor", pi, pj, dx, dy, r, dW, st.pos, st.vel, st.wall_vel, st.Vol, st.B,
                        nf, prm.DL, prm.DH, prm.mu, prm.h, nargout=2)
    st.tau_bottom, st.tau_top = float(np.ravel(tb)[0]), float(np.ravel(tt)[0])
    return dts


# ------------------------------------------------------------------------------------------------
# monitors (m:579-605) and the end-of-run L2 (postprocess:42)
# ------------------------------------------------------------------------------------------------
def binned_profile(y, u, y_min, y_max, n_bins):
    """compute_binned_profile_mean, m:579-590: discretize() bins are left-closed, the last one is
    closed on both sides, out-of-range values are dropped."""
    edges = np.linspace(y_min, y_max, n_bins + 1)
    y_mid = 0.5 * (edges[:-1] + edges[1:])
    y = np.asarray(y, dtype=np.float64)
    u = np.asarray(u, dtype=np.float64)
    b = np.searchsorted(edges, y, side="right") - 1
    b[y == edges[-1]] = n_bins - 1
    ok = (y >= edges[0]) & (y <= edges[-1])
    s = np.bincount(b[ok], weights=u[ok], minlength=n_bins)[:n_bins]
    c = np.bincount(b[ok], minlength=n_bins)[:n_bins]
    mean = s / np.maximum(c, 1)
    mean[c == 0] = np.nan
    return y_mid, mean


def mid_channel_profile(pos_f, u_x, DL, DH, mid_x, half_width, n_bins):  # m:592-605
    xw = np.mod(pos_f[:, 0], DL)
    d = np.abs(xw - mid_x)
    d = np.minimum(d, DL - d)
    sel = d <= half_width
    return binned_profile(pos_f[sel, 1], u_x[sel], 0.0, DH, n_bins)


def l2_error(prm: Params, pos_f, u_x):
    """Full-length L2 of the binned u(y) vs analytic (make_postprocess_data m:617-623, postprocess:42)."""
    n_bins = max(20, int(_matlab_round(prm.DH / prm.dp)))
    y_mid, u_mean = binned_profile(pos_f[:, 1], u_x, 0.0, prm.DH, n_bins)
    u_exact = prm.gravity_g / (2.0 * prm.nu) * y_mid * (prm.DH - y_mid)
    ok = ~np.isnan(u_mean)
    return math.sqrt(np.sum((u_mean[ok] - u_exact[ok]) ** 2) / max(np.sum(u_exact[ok] ** 2), np.finfo(float).eps))
