"""Builds tests/cpusim/_build/libsph_b200_cpusim.so: the product's CUDA sources compiled for the HOST against the
CPU functional simulator (include/cuda_runtime.h, cpusim_rt.cpp).  TEST INFRASTRUCTURE -- see the header comment of
include/cuda_runtime.h: nothing here is shipped or loaded by the product, and nothing measured with it is a result.

    python tests/cpusim/build.py [--force] [-DSPHB_FLAT_STAGING=1 ...]

The .cu files are copied into _build/src with three mechanical rewrites (the originals are never touched):
  * kernel<<<grid, block, smem, stream>>>(args)   ->  cpusim::launch(kernel, grid, block, smem, stream, args)
  * extern __shared__ ... smem_raw[];             ->  unsigned char *smem_raw = cpusim::dyn_smem();
  * the two inline-PTX SFU seeds of physmath.cuh  ->  cpusim::rcp_approx / cpusim::rsqrt_approx
slab.cu is left out (its callers need NCCL and torch streams).
"""
from __future__ import annotations

import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "sph-poiseuille-flow_b200", "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libsph_b200_cpusim.so")
DRIVER = os.path.join(OUT, "sph_poiseuille_cpusim")
MEXHOST = os.path.join(OUT, "libsphb_mex_stubhost_cpusim.so")
UNITS = {"misc.cu": [], "scan.cu": [], "neighbor.cu": ["-ffp-contract=off"], "sort.cu": ["-ffp-contract=off"],
         "ops_pairs.cu": [], "ops_csr.cu": [], "context.cu": [], "slab.cu": [], "slab_group.cu": [], "batch.cu": [], "api.cu": []}


def _matching(text: str, start: int, open_ch: str, close_ch: str) -> int:
    """Index of the bracket closing the one at `start`."""
    depth = 0
    for i in range(start, len(text)):
        if text[i] == open_ch:
            depth += 1
        elif text[i] == close_ch:
            depth -= 1
            if depth == 0:
                return i
    raise ValueError("unbalanced brackets")


def _split_top(s: str):
    parts, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur.strip())
            cur = ""
        else:
            cur += ch
    parts.append(cur.strip())
    return parts


def rewrite_launches(text: str) -> str:
    out, pos = "", 0
    while True:
        k = text.find("<<<", pos)
        if k < 0:
            return out + text[pos:]
        # kernel expression: identifier, optionally followed by template arguments <...>
        j = k
        if text[j - 1] == ">":
            depth = 0
            while True:
                j -= 1
                if text[j] == ">":
                    depth += 1
                elif text[j] == "<":
                    depth -= 1
                    if depth == 0:
                        break
        while j > 0 and (text[j - 1].isalnum() or text[j - 1] in "_:"):
            j -= 1
        kernel = text[j:k]
        e = text.index(">>>", k)
        cfg = _split_top(text[k + 3:e])
        while len(cfg) < 3:
            cfg.append("0")
        if len(cfg) < 4:
            cfg.append("nullptr")
        a0 = e + 3
        while text[a0] in " \n\t":
            a0 += 1
        assert text[a0] == "(", f"launch of {kernel}: argument list expected"
        a1 = _matching(text, a0, "(", ")")
        args = text[a0 + 1:a1].strip()
        call = f"cpusim::launch({kernel}, {cfg[0]}, {cfg[1]}, (size_t)({cfg[2]}), {cfg[3]}" + (f", {args})" if args else ")")
        out += text[pos:j] + call
        pos = a1 + 1


def rewrite(text: str) -> str:
    text = rewrite_launches(text)
    text = re.sub(r"extern\s+__shared__\s+__align__\(16\)\s+unsigned char smem_raw\[\];",
                  "unsigned char *smem_raw = cpusim::dyn_smem();", text)
    text = re.sub(r'asm\("rcp\.approx\.ftz\.f64 %0, %1;" : "=d"\((\w+)\) : "d"\((\w+)\)\);',
                  r"\1 = cpusim::rcp_approx(\2);", text)
    text = re.sub(r'asm\("rsqrt\.approx\.ftz\.f64 %0, %1;" : "=d"\((\w+)\) : "d"\((\w+)\)\);',
                  r"\1 = cpusim::rsqrt_approx(\2);", text)
    assert "<<<" not in text and "asm(" not in text and "extern __shared__" not in text
    return text


def build(force: bool = False, defines=()) -> str:
    src_out = os.path.join(OUT, "src", "sph-poiseuille-flow_b200", "csrc")
    inc_out = os.path.join(OUT, "src", "include")
    os.makedirs(src_out, exist_ok=True)
    os.makedirs(inc_out, exist_ok=True)
    stamp = os.path.join(OUT, "flags.txt")
    flags = " ".join(list(defines) + (["asan"] if os.environ.get("CPUSIM_ASAN") else []))
    if not os.path.exists(stamp) or open(stamp).read() != flags:
        force = True
    newest = 0.0
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cu", ".cuh")):
            newest = max(newest, os.path.getmtime(os.path.join(CSRC, f)))
            with open(os.path.join(CSRC, f)) as fh:
                txt = rewrite(fh.read())
            dst = os.path.join(src_out, f)
            if not os.path.exists(dst) or open(dst).read() != txt:
                open(dst, "w").write(txt)
    hdr = os.path.join(ROOT, "include", "sph_b200.h")
    newest = max(newest, os.path.getmtime(hdr), os.path.getmtime(os.path.join(HERE, "cpusim_rt.cpp")),
                 os.path.getmtime(os.path.join(HERE, "include", "cuda_runtime.h")), os.path.getmtime(__file__))
    open(os.path.join(inc_out, "sph_b200.h"), "w").write(open(hdr).read())
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= newest:
        return LIB
    san = ["-fsanitize=address", "-fno-omit-frame-pointer"] if os.environ.get("CPUSIM_ASAN") else []
    common = ["g++", "-std=c++17", "-O1", "-g", "-fPIC", "-fvisibility=hidden", "-fno-ipa-reference", "-fno-ipa-pure-const",
              "-Wno-unknown-pragmas", "-Wno-unused-function", "-Wno-attributes", "-U_FORTIFY_SOURCE", "-x", "c++",
              f"-I{os.path.join(HERE, 'include')}"] + san + list(defines)
    objs, procs = [], []
    for unit, extra in UNITS.items():
        obj = os.path.join(OUT, unit.replace(".cu", ".o"))
        objs.append(obj)
        procs.append((unit, subprocess.Popen(common + extra + ["-c", os.path.join(src_out, unit), "-o", obj],
                                             stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    rt = os.path.join(OUT, "cpusim_rt.o")
    procs.append(("cpusim_rt.cpp", subprocess.Popen(common + ["-c", os.path.join(HERE, "cpusim_rt.cpp"), "-o", rt],
                                                    stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for name, p in procs:
        log = p.communicate()[0]
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"--- {name}\n{log[:6000]}\n")
    if failed:
        raise RuntimeError("cpusim build failed")
    subprocess.run(["g++", "-shared", "-o", LIB] + san + objs + [rt, "-lm"], check=True)
    # the standalone C driver against the simulator library (tests/test_driver.py through SPHB_CPUSIM_DRIVER)
    drv = os.path.join(ROOT, "sph-poiseuille-flow_b200", "driver", "sph_poiseuille_main.c")
    subprocess.run(["gcc", "-O2", f"-I{os.path.join(ROOT, 'include')}", drv, "-o", DRIVER, f"-L{OUT}",
                    "-lsph_b200_cpusim", "-Wl,-rpath,$ORIGIN", "-lm"] + san, check=True)
    # the two MEX shims + the stub MATLAB host (oracle/stub) against the simulator library
    stub, mexdir = os.path.join(ROOT, "oracle", "stub"), os.path.join(ROOT, "sph-poiseuille-flow_b200", "mex")
    mobjs = []
    for src, sym in (("sph_neighbor_search_mex.c", "sphb_neighbor_mexFunction"),
                     ("sph_physics_shell_mex.c", "sphb_physics_mexFunction")):
        obj = os.path.join(OUT, src.replace(".c", ".o"))
        subprocess.run(["gcc", "-O2", "-fPIC", f"-I{stub}", f"-I{os.path.join(ROOT, 'include')}", f"-DmexFunction={sym}",
                        "-c", os.path.join(mexdir, src), "-o", obj] + san, check=True)
        mobjs.append(obj)
    hobj = os.path.join(OUT, "mexhost.o")
    subprocess.run(["gcc", "-O2", "-fPIC", "-Wno-use-after-free", f"-I{stub}", "-c", os.path.join(stub, "mexhost.c"),
                    "-o", hobj] + san, check=True)
    subprocess.run(["gcc", "-shared", "-o", MEXHOST] + san + mobjs + [hobj, f"-L{OUT}", "-lsph_b200_cpusim",
                                                                    "-Wl,-rpath,$ORIGIN", "-lm"], check=True)
    open(stamp, "w").write(flags)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, defines=[a for a in sys.argv[1:] if a.startswith("-D")]))
