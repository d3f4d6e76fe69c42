#!/usr/bin/env python
"""Source-line hot spots of one kernel from an ncu report captured with --import-source on.

    ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:k_force_fused > ff.csv
    python tools/ncu_source_lines.py ff.csv [top]

Per CUDA source line (inlined device functions are attributed to their own lines): share of the warp stall
samples, share of the executed warp instructions, shared-memory wavefronts (actual / ideal) and the dominant
stall reasons.  Reads the "cuda,sass" CSV of ncu's source page (rows with a line number are per-line totals).
"""
import csv
import sys


def main():
    path = sys.argv[1]
    top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    rows = list(csv.reader(open(path, errors="replace")))
    hdr = next(r for r in rows if r and r[0] == "Line No")
    col = {}
    for i, h in enumerate(hdr):
        col.setdefault(h, i)
    i_smp, i_ins = col["# Samples"], col["Instructions Executed"]
    i_w, i_wi = col.get("L1 Wavefronts Shared"), col.get("L1 Wavefronts Shared Ideal")
    stall_cols = [(h, i) for h, i in col.items() if h.startswith("stall_") and "Not Issued" not in h]
    lines = []
    fpath = ""
    for r in rows:
        if r and r[0] == "File Path":
            fpath = r[1].split("/")[-1]
        if not r or not r[0].isdigit() or len(r) <= i_ins:
            continue

        def f(i):
            try:
                return float(r[i])
            except (ValueError, IndexError, TypeError):
                return 0.0
        st = sorted(((f(i), h[6:]) for h, i in stall_cols), reverse=True)[:3]
        lines.append(dict(file=fpath, line=int(r[0]), src=r[1].strip(), smp=f(i_smp), ins=f(i_ins),
                          wf=f(i_w) if i_w else 0.0, wfi=f(i_wi) if i_wi else 0.0, stalls=st))
    tot_s = sum(x["smp"] for x in lines) or 1.0
    tot_i = sum(x["ins"] for x in lines) or 1.0
    tot_w, tot_wi = sum(x["wf"] for x in lines), sum(x["wfi"] for x in lines)
    print(f"total: {tot_s:.0f} stall samples, {tot_i:.0f} warp instructions, shared wavefronts {tot_w:.0f} "
          f"(ideal {tot_wi:.0f}, x{tot_w / max(tot_wi, 1):.2f})")
    print("samples%  instr%  smem wavefronts (x ideal)  top stalls                         file:line  source")
    for x in sorted(lines, key=lambda x: -x["smp"])[:top_n]:
        st = " ".join(f"{n}:{100 * v / max(x['smp'], 1):.0f}%" for v, n in x["stalls"] if v > 0)
        wf = f"{100 * x['wf'] / max(tot_w, 1):5.1f}% (x{x['wf'] / x['wfi']:.2f})" if x["wfi"] else "    -        "
        print(f"{100 * x['smp'] / tot_s:7.1f} {100 * x['ins'] / tot_i:7.1f}   {wf:<22} {st:<34} {x['file']}:{x['line']}  {x['src'][:90]}")


if __name__ == "__main__":
    main()
