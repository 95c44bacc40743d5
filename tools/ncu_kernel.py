"""Prints the pipe / stall / memory metrics of one kernel from `ncu --page raw --csv` output: ncu_kernel.py file.csv substring [index]"""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
kn = hdr.index('Kernel Name')
sel = [r for r in rows[2:] if sys.argv[2] in r[kn]]
r = sel[int(sys.argv[3]) if len(sys.argv) > 3 else 0]
pat = re.compile(r'^(Kernel Name|Grid Size|gpu__time_duration.sum|dram__bytes_(read|write).sum$|sm__pipe_\w+_cycles_active.avg.pct_of_peak_sustained_active|'
                 r'sm__inst_executed_pipe_\w+.avg.pct_of_peak_sustained_active|smsp__average_warps_issue_stalled_\w+_per_issue_active.ratio|'
                 r'smsp__issue_active.avg.pct|sm__warps_active.avg.per_cycle_active|smsp__inst_executed.sum$|sm__throughput.avg.pct|'
                 r'l1tex__data_pipe_lsu_wavefronts(_mem_shared)?.sum(.pct_of_peak_sustained_elapsed)?$|l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum$|'
                 r'l1tex__throughput.avg.pct_of_peak_sustained_active|lts__throughput.avg.pct|l1tex__t_sector_hit_rate.pct|lts__t_sector_hit_rate.pct|'
                 r'launch__registers_per_thread|launch__occupancy_limit_\w+|smsp__thread_inst_executed_per_inst_executed.ratio|l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum$|l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum$|dram__throughput.avg.pct_of_peak_sustained_elapsed|l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed|l1tex__data_pipe_lsu_wavefronts_mem_lgds.sum)')
for i, h in enumerate(hdr):
    if pat.match(h) and r[i] not in ('', 'n/a'):
        print(f"{h:88s} {r[i]:>18s} {units[i]}")
