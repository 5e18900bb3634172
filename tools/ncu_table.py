"""Compact per-kernel table from an `ncu --page raw --csv` export: python tools/ncu_table.py raw.csv [min_time_ms]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
H, data = rows[0], rows[2:]
tmin = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
K = [('gpu__time_duration.sum', 't_ms'), ('dram__bytes_read.sum', 'rdMB'), ('dram__bytes_write.sum', 'wrGB'),
     ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps%'), ('launch__registers_per_thread', 'regs'),
     ('smsp__inst_executed.sum', 'inst'), ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue%'),
     ('sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'fp64%'),
     ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smem_wf'), ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'conf'),
     ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'long'),
     ('smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'short'),
     ('smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'wait'),
     ('smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio', 'mio'),
     ('smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'math'),
     ('smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio', 'br'),
     ('smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'nsel'),
     ('smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio', 'noi'),
     ('smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio', 'slp'),
     ('smsp__average_warps_issue_stalled_membar_per_issue_active.ratio', 'mbar'),
     ('smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'lg'),
     ('smsp__average_warps_issue_stalled_drain_per_issue_active.ratio', 'drn'),
     ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram%'), ('l1tex__throughput.avg.pct_of_peak_sustained_active', 'l1%'),
     ('launch__occupancy_limit_shared_mem', 'oS'), ('launch__occupancy_limit_registers', 'oR')]
print('kernel'.ljust(28), ' '.join(n.rjust(8) for _, n in K))
for r in data:
    name = r[H.index('Kernel Name')].replace('void ', '')[:28]
    t = float(r[H.index('gpu__time_duration.sum')])
    if t < tmin: continue
    out = []
    for k, n in K:
        v = r[H.index(k)] if k in H else ''
        try:
            f = float(v.replace(',', ''))
            out.append((f'{f:8.3g}' if abs(f) < 1e5 else f'{f:8.2e}'))
        except ValueError:
            out.append(v[:8].rjust(8))
    print(name.ljust(28), ' '.join(out))
