"""sph_b200 -- Python host mirror of the reference's MEX call surface over the C ABI of libsph_b200.

The reference (KIYOYOZU/SPH-Poiseuille-Flow) is driven from MATLAB through two MEX files
(SPH_Poiseuille.m:38-41).  MATLAB is not available here, so this module mirrors that plugin surface
one-to-one in Python (same names, argument order, output order, error identifiers) on top of the
C-ABI library `libsph_b200.so` (include/sph_b200.h).  It contains no arithmetic: everything is done
by the hand-written sm_100a kernels behind the ABI, and there is no CPU fallback -- importing works
anywhere, but any compute call without a CUDA device raises SphError('SPHB:cuda').

    sph_neighbor_search_mex(pos, n_fluid, n_total, h, DL) -> (pair_i, pair_j, dx, dy, r, W, dW)
    sph_physics_shell_mex(mode, *inputs, nargout=k)       -> list of outputs, modes as in
        mex/sph_physics_mex.c:1753-1768

    Channel(params)  -- the device-resident context (sphb_ctx_*): upload, neighbour build, single
                        operators, reference steps, dual-criteria runs, monitors.

The package directory is named `sph-poiseuille-flow_b200` (not an identifier); load it with
`tests/conftest.py:load_product()` or `__graft_entry__.load_product()`, which register it as `sph_b200`.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from . import channel_setup  # noqa: F401  (host-side set-up mirror of SPH_Poiseuille.m S2/S3)
from .channel_setup import ChannelConfig  # noqa: F401

HERE = os.path.dirname(os.path.abspath(__file__))
# SPHB_LIB_VARIANT=<tag> loads libsph_b200_<tag>.so, an A/B build of the same sources with experimental -D
# switches (build.py, SPHB_BUILD_TAG); a missing variant fails loudly like a missing default build.
_VARIANT = os.environ.get("SPHB_LIB_VARIANT", "")
LIB_PATH = os.path.join(HERE, f"libsph_b200_{_VARIANT}.so" if _VARIANT else "libsph_b200.so")

_DP = C.POINTER(C.c_double)


class SphError(RuntimeError):
    """Raised for a non-zero return of the C ABI; `identifier` is the MATLAB-style error id."""

    def __init__(self, identifier: str, message: str):
        super().__init__(f"{identifier}: {message}")
        self.identifier = identifier
        self.message = message


class _Mat(C.Structure):
    _fields_ = [("p", _DP), ("m", C.c_size_t), ("n", C.c_size_t)]


class _Params(C.Structure):
    _fields_ = [("n_fluid", C.c_int), ("n_total", C.c_int),
                ("DL", C.c_double), ("DH", C.c_double), ("dp", C.c_double), ("h", C.c_double),
                ("rho0", C.c_double), ("p0", C.c_double), ("c_f", C.c_double), ("mu", C.c_double),
                ("gravity_g", C.c_double), ("inv_sigma0", C.c_double), ("transport_coeff", C.c_double),
                ("sort_interval", C.c_int), ("device", C.c_int), ("max_neighbors_hint", C.c_int)]


class _Status(C.Structure):
    _fields_ = [("t", C.c_double), ("dt", C.c_double), ("vmax", C.c_double), ("tau_bottom", C.c_double),
                ("tau_top", C.c_double), ("step", C.c_longlong), ("n_half_pairs", C.c_longlong),
                ("n_csr_entries", C.c_longlong), ("error_flags", C.c_int)]


class _Timing(C.Structure):
    _fields_ = [("neighbor_ms", C.c_float), ("density_ms", C.c_float), ("visc_transport_ms", C.c_float),
                ("verlet_ms", C.c_float), ("shear_ms", C.c_float), ("sort_ms", C.c_float), ("total_ms", C.c_float),
                ("launches", C.c_longlong)]


class _KProf(C.Structure):
    _fields_ = [("n", C.c_int), ("name", C.c_char_p * 24), ("ms", C.c_double * 24), ("calls", C.c_longlong * 24)]


_lib = None


def lib():
    """The loaded libsph_b200.so (built in-tree by build.py; never a silent fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SphError("SPHB:build", f"{LIB_PATH} is missing: run `python sph-poiseuille-flow_b200/build.py` "
                                         "(nvcc, sm_100a); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        L.sphb_last_error_id.restype = C.c_char_p
        L.sphb_last_error_msg.restype = C.c_char_p
        L.sphb_ctx_stream.restype = C.c_void_p
        L.sphb_ctx_stream.argtypes = [C.c_void_p]
        L.sphb_ctx_destroy.restype = None
        L.sphb_ctx_destroy.argtypes = [C.c_void_p]
        L.sphb_neighbor_release.restype = None
        L.sphb_neighbor_release.argtypes = [C.c_void_p]
        L.sphb_batch_destroy.restype = None
        L.sphb_batch_destroy.argtypes = [C.c_void_p]
        L.sphb_batch_launches.restype = C.c_longlong
        L.sphb_batch_launches.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _check(rc: int):
    if rc != 0:
        L = lib()
        raise SphError((L.sphb_last_error_id() or b"SPHB:error").decode(),
                       (L.sphb_last_error_msg() or b"").decode(errors="replace"))


def device_available() -> bool:
    return lib().sphb_device_check() == 0


def fp64_peak(iters: int = 4096, reps: int = 5) -> float:
    """Measured DFMA throughput of the current device in TFLOP/s (sphb_fp64_peak)."""
    out = C.c_double(0.0)
    _check(lib().sphb_fp64_peak(C.c_int(iters), C.c_int(reps), C.byref(out)))
    return out.value


# ------------------------------------------------------------------------------------------------
# marshalling: MATLAB values -> sphb_mat
# ------------------------------------------------------------------------------------------------
def _as_matrix(a) -> np.ndarray:
    arr = np.asarray(a, dtype=np.float64)
    if arr.ndim == 0:
        arr = arr.reshape(1, 1)
    elif arr.ndim == 1:
        arr = arr.reshape(-1, 1)      # MATLAB column vector
    return np.asfortranarray(arr)


def _mat(arr: np.ndarray) -> _Mat:
    return _Mat(arr.ctypes.data_as(_DP), arr.shape[0], arr.shape[1])


def _scalar(a) -> float:
    return float(np.asarray(a, dtype=np.float64).ravel()[0])   # mxGetScalar


def _int(a) -> int:
    return int(_scalar(a))                                      # (int)mxGetScalar


def _out(*shape) -> np.ndarray:
    return np.zeros(shape, order="F")


def _p(a: np.ndarray):
    return a.ctypes.data_as(_DP)


# ------------------------------------------------------------------------------------------------
# sph_neighbor_search_mex  (mex/sph_neighbor_search_mex.c:185-421)
# ------------------------------------------------------------------------------------------------
def sph_neighbor_search_mex(*args, nargout: int = 7):
    """[pair_i, pair_j, dx, dy, r, W, dW] = sph_neighbor_search_mex(pos, n_fluid, n_total, h, DL)"""
    if len(args) != 5:
        raise SphError("SPH:Neighbor:nrhs", "Expected 5 inputs.")
    if nargout != 7:
        raise SphError("SPH:Neighbor:nlhs", "Expected 7 outputs.")
    pos = _as_matrix(args[0])
    L = lib()
    handle = C.c_void_p()
    n_pairs = C.c_size_t(0)
    _check(L.sphb_neighbor_search(_mat(pos), C.c_int(_int(args[1])), C.c_int(_int(args[2])),
                                  C.c_double(_scalar(args[3])), C.c_double(_scalar(args[4])),
                                  C.byref(handle), C.byref(n_pairs)))
    outs = [np.zeros(n_pairs.value) for _ in range(7)]
    _check(L.sphb_neighbor_fetch(handle, *[_p(o) for o in outs]))
    return tuple(outs)


# ------------------------------------------------------------------------------------------------
# sph_physics_shell_mex  (mex/sph_physics_mex.c:1745-1772)
# ------------------------------------------------------------------------------------------------
def _nargs(cond: bool, ident: str, msg: str):
    if not cond:
        raise SphError(ident, msg)


def sph_physics_shell_mex(*args, nargout: int):
    """outputs = sph_physics_shell_mex(mode, ...) -- the eight modes of the reference's physics MEX."""
    _nargs(len(args) >= 1, "SPH:Physics:nrhs", "At least mode input is required.")
    _nargs(isinstance(args[0], str), "SPH:Physics:mode", "First input must be mode string.")
    mode, a = args[0], args[1:]
    nrhs = len(args)
    L = lib()
    d, ci = C.c_double, C.c_int

    def mats(*idx):
        arrs = [_as_matrix(a[k]) for k in idx]
        return arrs, [_mat(x) for x in arrs]

    if mode == "density_correction":
        _nargs(nrhs == 14, "SPH:Physics:density:nrhs", "density_correction expects 13 inputs after mode.")
        _nargs(nargout == 3, "SPH:Physics:density:nlhs", "density_correction expects 3 outputs.")
        keep, m = mats(0, 1, 2, 3, 4, 5, 6, 7)
        nt = max(_int(a[9]), 0)
        rho, Vol, B = _out(nt), _out(nt), _out(nt, 4)
        _check(L.sphb_density_correction(*m, ci(_int(a[8])), ci(_int(a[9])), d(_scalar(a[10])), d(_scalar(a[11])),
                                         d(_scalar(a[12])), _p(rho), _p(Vol), _p(B)))
        return [rho, Vol, B]
    if mode == "viscous_force":
        _nargs(nrhs in (16, 17), "SPH:Physics:viscous:nrhs", "viscous_force expects 15 inputs after mode.")
        _nargs(nargout == 1, "SPH:Physics:viscous:nlhs", "viscous_force expects 1 output.")
        keep, m = mats(0, 1, 2, 3, 4, 5, 6, 7, 8)
        keep2, m2 = mats(13, 14)
        nt = max(_int(a[12]), 0)
        force = _out(nt, 2)
        _check(L.sphb_viscous_force(*m, d(_scalar(a[9])), d(_scalar(a[10])), ci(_int(a[11])), ci(_int(a[12])),
                                    m2[0], m2[1], _p(force)))
        return [force]
    if mode == "transport_correction":
        _nargs(nrhs in (13, 14), "SPH:Physics:transport:nrhs",
               "transport_correction expects 12 or 13 inputs after mode.")
        _nargs(nargout == 1, "SPH:Physics:transport:nlhs", "transport_correction expects 1 output.")
        keep, m = mats(0, 1, 2, 3, 4, 5, 6, 7, 8)
        nt = max(_int(a[11]), 0)
        has = nrhs == 14
        out = _out(nt, 2)
        _check(L.sphb_transport_correction(*m, d(_scalar(a[9])), ci(_int(a[10])), ci(_int(a[11])), ci(1 if has else 0),
                                           d(_scalar(a[12]) if has else 0.0), _p(out)))
        return [out]
    if mode in ("integration_1st", "integration_verlet"):
        tag = "int1" if mode == "integration_1st" else "verlet"
        want = 5 if mode == "integration_1st" else 6
        _nargs(nrhs == 22, f"SPH:Physics:{tag}:nrhs", f"{mode} expects 21 inputs after mode.")
        _nargs(nargout == want, f"SPH:Physics:{tag}:nlhs", f"{mode} expects {want} outputs.")
        keep, m = mats(*range(14))
        keep2, m2 = mats(20)
        nt = max(_int(a[16]), 0)
        scal = (d(_scalar(a[14])), ci(_int(a[15])), ci(_int(a[16])), d(_scalar(a[17])), d(_scalar(a[18])),
                d(_scalar(a[19])))
        rho, p, drho = _out(nt), _out(nt), _out(nt)
        pos, vel, force = _out(nt, 2), _out(nt, 2), _out(nt, 2)
        if mode == "integration_1st":
            _check(L.sphb_integration_1st(*m, *scal, m2[0], _p(rho), _p(p), _p(pos), _p(force), _p(drho)))
            return [rho, p, pos, force, drho]
        _check(L.sphb_integration_verlet(*m, *scal, m2[0], _p(rho), _p(p), _p(pos), _p(vel), _p(drho), _p(force)))
        return [rho, p, pos, vel, drho, force]
    if mode == "integration_2nd":
        _nargs(nrhs == 15, "SPH:Physics:int2:nrhs", "integration_2nd expects 15 inputs after mode.")
        _nargs(nargout == 3, "SPH:Physics:int2:nlhs", "integration_2nd expects 3 outputs.")
        keep, m = mats(*range(10))
        keep2, m2 = mats(13)
        nt = max(_int(a[12]), 0)
        pos, drho, zeros = _out(nt, 2), _out(nt), _out(nt, 2)
        _check(L.sphb_integration_2nd(*m, d(_scalar(a[10])), ci(_int(a[11])), ci(_int(a[12])), m2[0], _p(pos),
                                      _p(drho), _p(zeros)))
        return [pos, drho, zeros]
    if mode == "advance_shell_step":
        _nargs(nrhs == 24, "SPH:Physics:advance:nrhs", "advance_shell_step expects 23 inputs after mode.")
        _nargs(nargout == 9, "SPH:Physics:advance:nlhs", "advance_shell_step expects 9 outputs.")
        keep, m = mats(*range(13))
        nt = max(_int(a[15]), 0)
        scal = (d(_scalar(a[13])), ci(_int(a[14])), ci(_int(a[15])), d(_scalar(a[16])), d(_scalar(a[17])),
                d(_scalar(a[18])), d(_scalar(a[19])), d(_scalar(a[20])), d(_scalar(a[21])), d(_scalar(a[22])))
        rho, p, drho, Vol = _out(nt), _out(nt), _out(nt), _out(nt)
        pos, vel, force, fp = _out(nt, 2), _out(nt, 2), _out(nt, 2), _out(nt, 2)
        B = _out(nt, 4)
        _check(L.sphb_advance_shell_step(*m, *scal, _p(rho), _p(p), _p(pos), _p(vel), _p(drho), _p(force), _p(fp),
                                         _p(Vol), _p(B)))
        return [rho, p, pos, vel, drho, force, fp, Vol, B]
    if mode == "wall_shear_monitor":
        _nargs(nrhs == 17, "SPH:Physics:wallshear:nrhs", "wall_shear_monitor expects 16 inputs after mode.")
        _nargs(nargout == 2, "SPH:Physics:wallshear:nlhs", "wall_shear_monitor expects 2 outputs.")
        keep, m = mats(*range(11))
        tb, tt = d(0.0), d(0.0)
        _check(L.sphb_wall_shear_monitor(*m, ci(_int(a[11])), d(_scalar(a[12])), d(_scalar(a[13])), d(_scalar(a[14])),
                                         d(_scalar(a[15])), C.byref(tb), C.byref(tt)))
        return [np.array([[tb.value]]), np.array([[tt.value]])]
    raise SphError("SPH:Physics:mode", "Unsupported mode.")


class MexBackend:
    """The same call surface as oracle.host.RefBackend / PortBackend, served by libsph_b200 (GPU)."""

    kind = "b200"

    def neighbor_search(self, pos, n_fluid, n_total, h, DL):
        return sph_neighbor_search_mex(pos, n_fluid, n_total, h, DL)

    def physics(self, mode, *args, nargout):
        out = sph_physics_shell_mex(mode, *args, nargout=nargout)
        return [o if (o.ndim == 2 and o.shape[1] > 1) else o.ravel(order="F") for o in out]


# ------------------------------------------------------------------------------------------------
# device-resident context
# ------------------------------------------------------------------------------------------------
@dataclass
class ChannelParams:
    n_fluid: int
    n_total: int
    DL: float
    DH: float
    dp: float
    h: float
    rho0: float
    p0: float
    c_f: float
    mu: float
    gravity_g: float
    inv_sigma0: float
    transport_coeff: float = 0.30
    sort_interval: int = 100
    device: int = -1
    max_neighbors_hint: int = 0


class Channel:
    """One periodic channel resident in HBM (sphb_ctx_*): SURVEY.md 8(a) rows N1-N6, P1-P9, S1."""

    def __init__(self, prm: ChannelParams):
        self.prm = prm
        self._ctx = C.c_void_p()
        cp = _Params(prm.n_fluid, prm.n_total, prm.DL, prm.DH, prm.dp, prm.h, prm.rho0, prm.p0, prm.c_f, prm.mu,
                     prm.gravity_g, prm.inv_sigma0, prm.transport_coeff, prm.sort_interval, prm.device,
                     prm.max_neighbors_hint)
        _check(lib().sphb_ctx_create(C.byref(cp), C.byref(self._ctx)))

    def close(self):
        if self._ctx:
            lib().sphb_ctx_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- transfers -------------------------------------------------------------------------------
    def upload(self, pos=None, vel=None, rho=None, drho_dt=None, mass=None, wall_vel=None):
        nt = self.prm.n_total
        keep = []

        def arg(a, cols):
            if a is None:
                return None
            arr = _as_matrix(a)
            if arr.shape != (nt, cols):
                raise SphError("SPHB:arg", f"expected an array of shape ({nt}, {cols}), got {arr.shape}")
            keep.append(arr)
            return _p(arr)

        _check(lib().sphb_ctx_upload(self._ctx, arg(pos, 2), arg(vel, 2), arg(rho, 1), arg(drho_dt, 1),
                                     arg(mass, 1), arg(wall_vel, 2)))

    def download(self, fields=("pos", "vel", "rho", "p", "drho_dt", "force", "force_prior", "Vol", "B")):
        nt = self.prm.n_total
        order = ("pos", "vel", "rho", "p", "drho_dt", "force", "force_prior", "Vol", "B")
        cols = dict(pos=2, vel=2, rho=1, p=1, drho_dt=1, force=2, force_prior=2, Vol=1, B=4)
        out = {k: (_out(nt, cols[k]) if cols[k] > 1 else _out(nt)) for k in fields}
        _check(lib().sphb_ctx_download(self._ctx, *[(_p(out[k]) if k in out else None) for k in order]))
        return out

    def download_into(self, out: dict):
        """Like download(), but into arrays of the caller (column-major float64 of the right shape, e.g.
        pinned host memory): no allocation, no extra copy."""
        order = ("pos", "vel", "rho", "p", "drho_dt", "force", "force_prior", "Vol", "B")
        cols = dict(pos=2, vel=2, rho=1, p=1, drho_dt=1, force=2, force_prior=2, Vol=1, B=4)
        nt = self.prm.n_total
        for k, a in out.items():
            want = (nt, cols[k]) if cols[k] > 1 else (nt,)
            if a.dtype != np.float64 or a.shape != want or not (a.flags.f_contiguous or a.ndim == 1):
                raise SphError("SPHB:arg", f"download_into: {k} must be a column-major float64 array of shape {want}")
        _check(lib().sphb_ctx_download(self._ctx, *[(_p(out[k]) if k in out else None) for k in order]))
        return out

    def set_time(self, t: float, step: int):
        _check(lib().sphb_ctx_set_time(self._ctx, C.c_double(t), C.c_longlong(step)))

    # -- neighbour list --------------------------------------------------------------------------
    def neighbor_build(self):
        _check(lib().sphb_ctx_neighbor_build(self._ctx))

    def neighbor_count(self):
        a, b = C.c_size_t(0), C.c_size_t(0)
        _check(lib().sphb_ctx_neighbor_count(self._ctx, C.byref(a), C.byref(b)))
        return a.value, b.value

    def export_pairs(self):
        n, _ = self.neighbor_count()
        outs = [np.zeros(n) for _ in range(7)]
        _check(lib().sphb_ctx_neighbor_export_pairs(self._ctx, *[_p(o) for o in outs]))
        return tuple(outs)

    def export_csr(self):
        _, n = self.neighbor_count()
        row_ptr = np.zeros(self.prm.n_fluid + 1, dtype=np.int32)
        col = np.zeros(max(n, 1), dtype=np.int32)
        _check(lib().sphb_ctx_neighbor_export_csr(self._ctx, row_ptr.ctypes.data_as(C.POINTER(C.c_int)),
                                                  col.ctypes.data_as(C.POINTER(C.c_int))))
        return row_ptr, col[:n]

    # -- single operators ------------------------------------------------------------------------
    def density_correction(self):
        _check(lib().sphb_ctx_density_correction(self._ctx))

    def eos(self):
        _check(lib().sphb_ctx_eos(self._ctx))

    def viscous_force(self, add_gravity=True):
        _check(lib().sphb_ctx_viscous_force(self._ctx, C.c_int(1 if add_gravity else 0)))

    def transport_correction(self):
        _check(lib().sphb_ctx_transport_correction(self._ctx))

    def integration_verlet(self, dt: float = 0.0):
        _check(lib().sphb_ctx_integration_verlet(self._ctx, C.c_double(dt)))

    def wall_shear(self):
        tb, tt = C.c_double(0.0), C.c_double(0.0)
        _check(lib().sphb_ctx_wall_shear(self._ctx, C.byref(tb), C.byref(tt)))
        return tb.value, tt.value

    def periodic_wrap(self):
        _check(lib().sphb_ctx_periodic_wrap(self._ctx))

    def sort_by_cell(self):
        _check(lib().sphb_ctx_sort_by_cell(self._ctx))

    # -- stepping --------------------------------------------------------------------------------
    def initialise(self):
        """S5 of the script (m:165-173): first list, density + KGC, EOS."""
        self.neighbor_build()
        self.density_correction()
        self.eos()

    def step(self, remain: float = float("inf")):
        _check(lib().sphb_ctx_step(self._ctx, C.c_double(remain)))

    def step_uploaded(self, remain: float = float("inf")):
        """One reference step on freshly uploaded state: list (if stale) + shear monitor + physics; no trailing
        list rebuild (sphb_ctx_step_uploaded)."""
        _check(lib().sphb_ctx_step_uploaded(self._ctx, C.c_double(remain)))

    def step_host(self, state: dict, out: dict, remain: float = float("inf")):
        """One reference step on host arrays in ONE call (sphb_ctx_step_host): `state` holds pos, vel, rho, drho_dt,
        `out` any of pos, vel, rho, p, drho_dt, force (column-major float64, ideally pinned); transfers overlap the list
        build.  Bit-identical to upload + step_uploaded + download_into."""
        nt = self.prm.n_total
        cols = dict(pos=2, vel=2, rho=1, p=1, drho_dt=1, force=2)
        for d in (state, out):
            for k, a in d.items():
                want = (nt, cols[k]) if cols[k] > 1 else (nt,)
                if a.dtype != np.float64 or a.shape != want or not (a.flags.f_contiguous or a.ndim == 1):
                    raise SphError("SPHB:arg", f"step_host: {k} must be a column-major float64 array of shape {want}")
        _check(lib().sphb_ctx_step_host(
            self._ctx, C.c_double(remain), _p(state["pos"]), _p(state["vel"]), _p(state["rho"]), _p(state["drho_dt"]),
            *[(_p(out[k]) if k in out else None) for k in ("pos", "vel", "rho", "p", "drho_dt", "force")]))
        return out

    def run_until(self, t_target: float, t_end: float, max_steps: int = 0) -> int:
        n = C.c_longlong(0)
        _check(lib().sphb_ctx_run_until(self._ctx, C.c_double(t_target), C.c_double(t_end), C.c_longlong(max_steps),
                                        C.byref(n)))
        return n.value

    def run_dual(self, n_outer: int, n_inner: int):
        _check(lib().sphb_ctx_run_dual(self._ctx, C.c_int(n_outer), C.c_int(n_inner)))

    def run_dual_auto(self, n_outer: int, max_inner: int, remain: float = float("inf")):
        """Dual-criteria schedule with Dt, dt and the substep count decided on the device (sphb_ctx_run_dual_auto)."""
        _check(lib().sphb_ctx_run_dual_auto(self._ctx, C.c_int(n_outer), C.c_int(max_inner), C.c_double(remain)))

    def dual_counts(self) -> dict:
        o, s, d = C.c_longlong(0), C.c_longlong(0), C.c_double(0.0)
        _check(lib().sphb_ctx_dual_counts(self._ctx, C.byref(o), C.byref(s), C.byref(d)))
        return {"outer_steps": o.value, "sub_steps": s.value, "last_Dt": d.value}

    def status(self) -> dict:
        s = _Status()
        _check(lib().sphb_ctx_status(self._ctx, C.byref(s)))
        return {k: getattr(s, k) for k, _ in _Status._fields_}

    def profile(self, n_bins: int, x0: float = -1.0, half_width: float = 0.0):
        u, c = np.zeros(n_bins), np.zeros(n_bins)
        _check(lib().sphb_ctx_profile(self._ctx, C.c_double(x0), C.c_double(half_width), C.c_int(n_bins), _p(u), _p(c)))
        return u, c

    def enable_timing(self, on=True):
        _check(lib().sphb_ctx_enable_timing(self._ctx, C.c_int(1 if on else 0)))

    def timing(self) -> dict:
        t = _Timing()
        _check(lib().sphb_ctx_timing(self._ctx, C.byref(t)))
        return {k: getattr(t, k) for k, _ in _Timing._fields_}

    def sync(self):
        _check(lib().sphb_ctx_sync(self._ctx))

    def kernel_profile(self, enable: int = -1, reset: bool = False) -> dict:
        """Per-kernel device time (CUDA events on the context stream): {name: (ms_total, calls)}."""
        k = _KProf()
        _check(lib().sphb_ctx_kernel_profile(self._ctx, C.c_int(enable), C.c_int(1 if reset else 0), C.byref(k)))
        return {k.name[i].decode(): (k.ms[i], k.calls[i]) for i in range(k.n)}

    @property
    def stream(self) -> int:
        return int(lib().sphb_ctx_stream(self._ctx) or 0)


class Ensemble:
    """Independent channels stepped together (BASELINE.json configs[3]): one device-resident context per
    channel, enqueued round-robin by sphb_ensemble_run_dual so their kernels overlap on the device."""

    def __init__(self, channels):
        self.channels = list(channels)
        self._arr = (C.c_void_p * len(self.channels))(*[ch._ctx for ch in self.channels])

    def run_dual(self, n_outer: int, n_inner: int = 1):
        _check(lib().sphb_ensemble_run_dual(self._arr, C.c_int(len(self.channels)), C.c_int(n_outer), C.c_int(n_inner)))

    def sync(self):
        _check(lib().sphb_ensemble_sync(self._arr, C.c_int(len(self.channels))))

    def close(self):
        for ch in self.channels:
            ch.close()


class Batch:
    """The channels of an ensemble stepped by BATCHED launches (sphb_batch_*): every kernel of the step is
    launched once for all channels (grid = the channel's own grid x channels), each channel with its own
    parameters, dt and clock.  Same kernel bodies and per-channel data as Channel.run_dual.  The channels must be
    initialised (first list build done), share n_fluid / n_total / h / DL / sort_interval and their cell grid,
    and live on one device; while the batch exists they launch on its stream."""

    def __init__(self, channels):
        self.channels = list(channels)
        self._arr = (C.c_void_p * len(self.channels))(*[ch._ctx for ch in self.channels])
        self._b = C.c_void_p()
        _check(lib().sphb_batch_create(self._arr, C.c_int(len(self.channels)), C.byref(self._b)))

    def run_dual(self, n_outer: int, n_inner: int = 1):
        _check(lib().sphb_batch_run_dual(self._b, C.c_int(n_outer), C.c_int(n_inner)))

    def refresh(self):
        """After a member's buffers or list capacity changed outside the batch (position upload, sort_by_cell,
        run_until)."""
        _check(lib().sphb_batch_refresh(self._b))

    def sync(self):
        _check(lib().sphb_batch_sync(self._b))

    @property
    def launches(self) -> int:
        return int(lib().sphb_batch_launches(self._b))

    def error_flags(self) -> int:
        f = C.c_int(0)
        _check(lib().sphb_batch_error_flags(self._b, C.byref(f)))
        return f.value

    def close(self, channels: bool = False):
        """Destroys the batch (its members get their own streams back); with channels=True the members too."""
        if self._b:
            lib().sphb_batch_destroy(self._b)
            self._b = C.c_void_p()
        if channels:
            for ch in self.channels:
                ch.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def channel_from_config(cfg, particles, **overrides) -> Channel:
    """Channel for a channel_setup.ChannelConfig (after .derive()) and its Particles, state uploaded."""
    kw = dict(n_fluid=particles.n_fluid, n_total=particles.n_total, DL=cfg.DL, DH=cfg.DH, dp=cfg.dp, h=cfg.h,
              rho0=cfg.rho0, p0=cfg.p0, c_f=cfg.c_f, mu=cfg.mu, gravity_g=cfg.gravity_g, inv_sigma0=cfg.inv_sigma0,
              transport_coeff=cfg.transport_coeff, sort_interval=cfg.sort_interval)
    kw.update(overrides)
    ch = Channel(ChannelParams(**kw))
    ch.upload(pos=particles.pos, vel=particles.vel, rho=particles.rho, drho_dt=particles.drho_dt,
              mass=particles.mass, wall_vel=particles.wall_vel)
    return ch


def channel_from_oracle_params(prm, n_fluid: int, n_total: int, **overrides) -> Channel:
    """Build a Channel from an oracle.host.Params-like object (same field names as the script's cfg)."""
    kw = dict(n_fluid=n_fluid, n_total=n_total, DL=prm.DL, DH=prm.DH, dp=prm.dp, h=prm.h, rho0=prm.rho0, p0=prm.p0,
              c_f=prm.c_f, mu=prm.mu, gravity_g=prm.gravity_g, inv_sigma0=prm.inv_sigma0,
              transport_coeff=prm.transport_coeff, sort_interval=prm.sort_interval)
    kw.update(overrides)
    return Channel(ChannelParams(**kw))
