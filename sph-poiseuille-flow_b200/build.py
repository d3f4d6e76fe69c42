"""Builds libsph_b200.so (sm_100a only) and the MEX shims, in-tree.

    python sph-poiseuille-flow_b200/build.py [--force] [--verbose]

Every CUDA translation unit is compiled with
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3
neighbor.cu and sort.cu additionally with -fmad=false (their floating-point expressions decide cell
membership and pair acceptance and must round exactly like the reference's, see neighbor.cu).
No Triton, no CUB/Thrust, no multi-arch fat binary.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
# A/B experiments: SPHB_BUILD_TAG=<tag> builds the library with the experimental -D switches below into
# libsph_b200_<tag>.so (objects in build_<tag>/) beside the default build; SPHB_LIB_VARIANT=<tag> makes the
# Python mirror load it.  Tagged builds skip the MEX shims and the driver.
TAG = os.environ.get("SPHB_BUILD_TAG", "")
OBJ = os.path.join(HERE, "build_" + TAG if TAG else "build")
LIB = os.path.join(HERE, f"libsph_b200_{TAG}.so" if TAG else "libsph_b200.so")
MEXHOST_LIB = os.path.join(HERE, "libsphb_mex_stubhost.so")
DRIVER = os.path.join(HERE, "sph_poiseuille_b200")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-std=c++17", "-O3", "-lineinfo", "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wall,-Wno-unused-function",
          "-Xptxas", "-v" if os.environ.get("SPHB_PTXAS_V") else "-O3"]
# experimental code paths, off unless asked for (each documented where it is defined)
for _flag in ("SPHB_FLAT_STAGING", "SPHB_STAGE_UNROLL"):
    if os.environ.get(_flag):
        COMMON.append(f"-D{_flag}={os.environ[_flag]}")
for _d in os.environ.get("SPHB_DEFINES", "").split():   # e.g. SPHB_DEFINES="SPHB_X=1 SPHB_Y=2"
    COMMON.append("-D" + _d)
UNITS = {
    "misc.cu": [],
    "scan.cu": [],
    "neighbor.cu": ["-fmad=false"],
    "sort.cu": ["-fmad=false"],
    "ops_pairs.cu": [],
    "ops_csr.cu": [],
    "context.cu": [],
    "slab.cu": [],
    "slab_group.cu": [],
    "batch.cu": [],
    "api.cu": [],
}
HOST_CC = "/usr/bin/gcc"   # the image's $CC wrapper lacks libgomp.spec; shims/driver are plain C


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libsph_b200 cannot be built (there is no CPU fallback)")


def _newer(target: str, deps) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def _run(cmd, verbose):
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("build step failed: " + " ".join(cmd))
    if verbose and (r.stdout or r.stderr):
        print(r.stdout + r.stderr)
    return r


def build_library(force=False, verbose=False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    # objects built with other flags (an experimental -D switch toggled) are stale
    stamp, flags = os.path.join(OBJ, "flags.txt"), " ".join(ARCH + COMMON)
    if not os.path.exists(stamp) or open(stamp).read() != flags:
        force = force or any(f.endswith(".o") for f in os.listdir(OBJ))
        open(stamp, "w").write(flags)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(ROOT, "include", "sph_b200.h"))
    objs = []
    for unit, extra in UNITS.items():
        src = os.path.join(CSRC, unit)
        obj = os.path.join(OBJ, unit.replace(".cu", ".o"))
        objs.append(obj)
        if not force and _newer(obj, [src] + headers):
            continue
        _run([nvcc] + ARCH + COMMON + extra + ["-c", src, "-o", obj], verbose)
    if force or not _newer(LIB, objs):
        _run([nvcc] + ARCH + ["-shared", "-o", LIB] + objs, verbose)
    return LIB


def build_mex_stubhost(force=False, verbose=False) -> str:
    """The MEX shims compiled against oracle/stub/mex.h + the fake MATLAB host (tests only; a real
    MATLAB build uses `mex` with MATLAB's own mex.h)."""
    mexdir = os.path.join(HERE, "mex")
    stub = os.path.join(ROOT, "oracle", "stub")
    srcs = [os.path.join(mexdir, "sph_neighbor_search_mex.c"), os.path.join(mexdir, "sph_physics_shell_mex.c")]
    if not all(os.path.exists(s) for s in srcs):
        return ""
    deps = srcs + [os.path.join(stub, "mex.h"), os.path.join(stub, "mexhost.c"), LIB]
    if not force and _newer(MEXHOST_LIB, deps):
        return MEXHOST_LIB
    os.makedirs(OBJ, exist_ok=True)
    objs = []
    for src, sym in zip(srcs, ("sphb_neighbor_mexFunction", "sphb_physics_mexFunction")):
        obj = os.path.join(OBJ, os.path.basename(src).replace(".c", ".o"))
        _run([HOST_CC, "-O2", "-fPIC", "-Wall", f"-I{stub}", f"-I{os.path.join(ROOT, 'include')}",
              f"-DmexFunction={sym}", "-c", src, "-o", obj], verbose)
        objs.append(obj)
    host_obj = os.path.join(OBJ, "mexhost.o")
    _run([HOST_CC, "-O2", "-fPIC", "-Wno-use-after-free", f"-I{stub}", "-c", os.path.join(stub, "mexhost.c"), "-o", host_obj],
         verbose)
    _run([HOST_CC, "-shared", "-o", MEXHOST_LIB] + objs + [host_obj, f"-L{HERE}", "-lsph_b200",
                                                           "-Wl,-rpath,$ORIGIN", "-lm"], verbose)
    return MEXHOST_LIB


def build_driver(force=False, verbose=False) -> str:
    src = os.path.join(HERE, "driver", "sph_poiseuille_main.c")
    if not os.path.exists(src):
        return ""
    if not force and _newer(DRIVER, [src, LIB]):
        return DRIVER
    _run([HOST_CC, "-O2", "-Wall", f"-I{os.path.join(ROOT, 'include')}", src, "-o", DRIVER, f"-L{HERE}", "-lsph_b200",
          "-Wl,-rpath,$ORIGIN", "-lm"], verbose)
    return DRIVER


def build_all(force=False, verbose=False):
    out = [build_library(force, verbose)]
    if TAG:
        return out
    out.append(build_mex_stubhost(force, verbose))
    out.append(build_driver(force, verbose))
    return [o for o in out if o]


if __name__ == "__main__":
    res = build_all(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print("built:", *res, sep="\n  ")
