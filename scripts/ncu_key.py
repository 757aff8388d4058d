"""Key metrics of every kernel in an ncu report (raw page).  usage: python scripts/ncu_key.py <report> [extra-substr ...]"""
import csv, io, subprocess, sys
KEYS = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__registers_per_thread ', 'launch__waves_per_multiprocessor',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum ', 'smsp__cycles_active.avg ', 'sm__cycles_elapsed.max ',
        'dram__bytes_read.sum ', 'dram__bytes_write.sum ', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
keys = KEYS + sys.argv[2:]
for r in rows[2:]:
    print("=" * 100)
    for h, u, v in zip(hdr, units, r):
        hh = h + " "
        if any(k in hh for k in keys):
            if 'issue_stalled' in h:
                try:
                    if float(v) < 0.15: continue
                except ValueError: pass
            print("%-95s %-12s %s" % (h, u, v))
