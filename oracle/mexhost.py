"""oracle/mexhost.py -- TEST INFRASTRUCTURE, not product code.

ctypes front end of the fake MATLAB host (oracle/stub/mexhost.c).  It lets a test call any
``mexFunction``-style gateway with numpy arrays the way ``SPH_Poiseuille.m`` calls
``feval(mex_name, ...)`` (reference call sites SPH_Poiseuille.m:167,169,366-432):

    lib  = MexLibrary(path_to_so)
    outs = lib.call("ref_neighbor_mexFunction", 7, pos, n_fluid, n_total, h, DL)

Arguments: ``str`` -> char array, python scalars -> 1x1 double, numpy arrays -> column-major
double matrices (1-D arrays become n x 1 columns, like MATLAB column vectors).  A MEX error
(``mexErrMsgIdAndTxt``) is raised as :class:`MexError` carrying the MATLAB error identifier.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm may import this.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np


class MexError(RuntimeError):
    """A MATLAB-style error raised by MEX code: ``identifier`` is e.g. 'SPH:Neighbor:count'."""

    def __init__(self, identifier: str, message: str):
        super().__init__(f"{identifier}: {message}")
        self.identifier = identifier
        self.message = message


class _Arg(C.Structure):
    _fields_ = [("is_char", C.c_int), ("str", C.c_char_p), ("m", C.c_size_t), ("n", C.c_size_t),
                ("data", C.c_void_p)]


class _Out(C.Structure):
    _fields_ = [("m", C.c_size_t), ("n", C.c_size_t), ("data", C.POINTER(C.c_double))]


class MexLibrary:
    """A shared library that links oracle/stub/mexhost.c and exports mexFunction-style gateways."""

    def __init__(self, path: str):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.path = path
        self.lib = C.CDLL(path, mode=C.RTLD_LOCAL)
        self.lib.mexhost_call.restype = C.c_int
        self.lib.mexhost_call.argtypes = [C.c_void_p, C.c_int, C.POINTER(_Out), C.c_int, C.POINTER(_Arg),
                                          C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
        self.lib.mexhost_free_outputs.restype = None
        self.lib.mexhost_free_outputs.argtypes = [C.c_int, C.POINTER(_Out)]

    def has(self, symbol: str) -> bool:
        return hasattr(self.lib, symbol)

    def call(self, symbol: str, nlhs: int, *args):
        fn = C.cast(getattr(self.lib, symbol), C.c_void_p)
        keep = []
        cargs = (_Arg * max(len(args), 1))()
        for k, a in enumerate(args):
            if isinstance(a, str):
                b = a.encode()
                keep.append(b)
                cargs[k] = _Arg(1, b, 1, len(b), None)
                continue
            arr = np.asarray(a, dtype=np.float64)
            if arr.ndim == 0:
                arr = arr.reshape(1, 1)
            elif arr.ndim == 1:
                arr = arr.reshape(-1, 1)
            arr = np.asfortranarray(arr)
            keep.append(arr)
            cargs[k] = _Arg(0, None, arr.shape[0], arr.shape[1], arr.ctypes.data)
        outs = (_Out * max(nlhs, 1))()
        errid = C.create_string_buffer(256)
        errmsg = C.create_string_buffer(1024)
        rc = self.lib.mexhost_call(fn, nlhs, outs, len(args), cargs, errid, 256, errmsg, 1024)
        if rc == 1:
            raise MexError(errid.value.decode(), errmsg.value.decode(errors="replace"))
        if rc != 0:
            raise RuntimeError(f"mexhost_call usage error rc={rc}")
        results = []
        for k in range(nlhs):
            m, n = outs[k].m, outs[k].n
            if outs[k].data and m * n > 0:
                flat = np.ctypeslib.as_array(outs[k].data, shape=(m * n,)).copy()
            else:
                flat = np.zeros(m * n)
            results.append(flat.reshape((m, n), order="F"))
        self.lib.mexhost_free_outputs(nlhs, outs)
        return results
