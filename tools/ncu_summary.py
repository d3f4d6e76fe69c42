"""Summarise an .ncu-rep (read here, no GPU): python tools/ncu_summary.py report.ncu-rep"""
import csv, subprocess, sys
rows = list(csv.reader(subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hdr = rows[0]
keys = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.sum', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_active.avg', 'sm__cycles_elapsed.max']
for r in rows[2:]:
    print('=====', r[hdr.index('Kernel Name')][:70])
    for i, h in enumerate(hdr):
        stall = 'smsp__average_warp' in h and 'issue_stalled' in h and h.endswith('.ratio') and 'not_issued' not in h
        if stall or h in keys:
            try:
                v = float(r[i].replace(',', ''))
            except ValueError:
                continue
            if not stall or v > 0.15:
                print('   %-90s %s %s' % (h.replace('smsp__average_warps_issue_stalled_', 'stall:').replace('_per_issue_active.ratio', ''), r[i], rows[1][i]))
