#!/usr/bin/env python
"""Per-kernel opcode histograms of two `cuobjdump -sass` dumps, compared.

    cuobjdump -sass old/libsph_b200.so > old.sass; cuobjdump -sass new/libsph_b200.so > new.sass
    python tools/sass_hist_diff.py old.sass new.sass [name-filter]

Used to check that a source refactor (e.g. a kernel body moved into a __device__ function shared by two
__global__ wrappers) left the machine code of the existing kernels alone without a GPU at hand: registers,
addresses and encodings are ignored; interchangeable forms ptxas alternates between for pipe balancing
(MOV / IMAD.MOV.U32, IADD3 / VIADD / IMAD.IADD, SHF.L / IMAD.SHL) are folded together.  Prints one line per
kernel present in both dumps: IDENTICAL-HIST or the opcodes whose counts differ (old, new), then the kernels
that exist in only one of them.
"""
import collections
import re
import sys

CANON = {"IMAD.MOV.U32": "MOV", "IMAD.MOV": "MOV", "VIADD": "IADD", "IADD3": "IADD", "IMAD.IADD": "IADD",
         "IMAD.SHL.U32": "SHL", "SHF.L.U32": "SHL"}


def parse(path):
    fn, out = None, {}
    for line in open(path, errors="replace"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            out[fn] = []
            continue
        if fn is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if m:
            out[fn].append(re.sub(r"\s+", " ", m.group(1)))
    return out


def hist(instrs):
    h = collections.Counter()
    for x in instrs:
        t = x.split()
        op = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
        h[CANON.get(op, op)] += 1
    return h


def short(name):
    m = re.search(r"\d+(k_\w+?|scan_\w+?)(?:I|E)", name)
    return m.group(1) if m else name[:60]


def main():
    a, b = parse(sys.argv[1]), parse(sys.argv[2])
    flt = sys.argv[3] if len(sys.argv) > 3 else ""
    same = 0
    for fn in sorted(a):
        if fn not in b or flt not in fn:
            continue
        ha, hb = hist(a[fn]), hist(b[fn])
        d = {k: (ha[k], hb[k]) for k in sorted(set(ha) | set(hb)) if ha[k] != hb[k]}
        same += not d
        print(short(fn).ljust(28), "IDENTICAL-HIST" if not d else d)
    print(f"identical: {same}; only in old: {sorted(short(f) for f in a if f not in b)}; "
          f"only in new: {sorted(short(f) for f in b if f not in a)}")


if __name__ == "__main__":
    main()
