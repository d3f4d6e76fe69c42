"""Hot SASS lines of one kernel from an .ncu-rep: python tools/ncu_hot.py report.ncu-rep kernel_regex [top]"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--kernel-name", "regex:" + sys.argv[2]],
                     capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) == len(hdr) and r[0].startswith("0x")]
tot = sum(int(r[col["# Samples"]]) for r in body)
tot_inst = sum(int(r[col["Instructions Executed"]]) for r in body)
print("samples", tot, "warp instructions", tot_inst, "sass lines", len(body))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(r[col[h]]) for r in body) for h in stalls}
print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > tot * 0.01})
# opcode mix
mix = {}
for r in body:
    op = r[col["Source"]].split()[0]
    if op.startswith("@"):
        op = r[col["Source"]].split()[1]
    op = op.split(".")[0]
    mix[op] = mix.get(op, 0) + int(r[col["Instructions Executed"]])
print("opcode mix:", ", ".join(f"{k}={v/tot_inst*100:.1f}%" for k, v in sorted(mix.items(), key=lambda kv: -kv[1])[:18]))
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
idx = sorted(range(len(body)), key=lambda k: -int(body[k][col["# Samples"]]))[:top]
for k in sorted(idx):
    r = body[k]
    st = {h[6:]: int(r[col[h]]) for h in stalls if int(r[col[h]]) > 0}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{k:5d} {int(r[col['# Samples']])/tot*100:5.2f}% ex={r[col['Instructions Executed']]:>8} thr={r[col['Avg. Threads Executed']]:>5} {r[col['Source']].strip()[:70]:70s} {st}")
