"""Condense an .ncu-rep (read on the CPU box) into the per-kernel JSON summary kept under profiles/.

    python tools/ncu_to_profile.py gpurun_out/prof.ncu-rep profiles/rN_vM_ncu_summary.json

For every captured kernel: duration, DRAM bytes (-> roofline.traffic in bench.py), FP64 / LSU pipe
utilisation, occupancy, issue rate, instruction count and the leading stall reasons."""
import csv
import json
import subprocess
import sys

rows = list(csv.reader(subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True,
                                      text=True).stdout.splitlines()))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}


def num(r, key, scale=1.0):
    try:
        return float(r[col[key]].replace(",", "")) * scale
    except (KeyError, ValueError):
        return None


def bytes_of(r, key):
    v = num(r, key)
    if v is None:
        return None
    u = units[col[key]].lower()
    return v * {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1.0)


out = {}
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    short = name.split("(")[0].split("::")[-1].split("<")[0].strip()
    stalls = {h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): num(r, h)
              for h in hdr if "issue_stalled" in h and h.endswith("_per_issue_active.ratio") and "not_issued" not in h}
    top = dict(sorted(((k, round(v, 3)) for k, v in stalls.items() if v), key=lambda kv: -kv[1])[:5])
    rd, wr = bytes_of(r, "dram__bytes_read.sum"), bytes_of(r, "dram__bytes_write.sum")
    rec = {
        "duration_us": num(r, "gpu__time_duration.sum"),
        "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes": (rd or 0) + (wr or 0),
        "dram_throughput_pct": num(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "fp64_pipe_pct": num(r, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
        "lsu_wavefronts_pct": num(r, "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
        "issue_active_pct": num(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "warps_active_pct": num(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
        "registers": num(r, "launch__registers_per_thread"), "grid": num(r, "launch__grid_size"),
        "dyn_smem_kb": num(r, "launch__shared_mem_per_block_dynamic"),
        "warp_instructions": num(r, "smsp__inst_executed.sum"),
        # DADD + DMUL + DFMA thread instructions of the launch (per-cycle sums x elapsed SMSP cycles): numerator of
        # the fp64 roofline in bench.py (peak = 64 per cycle and SM = sm__sass_thread_inst_executed_op_dfma_pred_on
        # .sum.peak_sustained / 148)
        "fp64_thread_instructions": (lambda c, ops: c * sum(ops) if c and all(o is not None for o in ops) else None)(
            num(r, "smsp__cycles_elapsed.avg"),
            [num(r, f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum.per_cycle_elapsed") for op in ("dadd", "dmul", "dfma")]),
        "fp64_peak_thread_inst_per_cycle": num(r, "sm__sass_thread_inst_executed_op_dfma_pred_on.sum.peak_sustained"),
        "threads_per_instruction": num(r, "smsp__thread_inst_executed_per_inst_executed.ratio"),
        "l1_hit_pct": num(r, "l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": num(r, "lts__t_sector_hit_rate.pct"),
        "stalls_per_issue": top,
    }
    out.setdefault(short, []).append(rec)
json.dump(out, open(sys.argv[2], "w"), indent=1)
for k, v in out.items():
    r = v[0]
    print(f"{k:28s} {r['duration_us']:8.1f} us  dram {r['dram_bytes']/1e6:7.1f} MB  fp64 {r['fp64_pipe_pct']:5.1f}%  "
          f"lsu {r['lsu_wavefronts_pct']:5.1f}%  issue {r['issue_active_pct']:5.1f}%  warps {r['warps_active_pct']:5.1f}%")
