import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
H = rows[0]
want = ['gpu__time_duration.sum', 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_alu.avg.pct', 'smsp__issue_active.avg.pct',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'launch__occupancy_limit_registers', 'launch__registers_per_thread', 'smsp__inst_executed.sum', 'smsp__average_warps_issue_stalled', 'launch__grid_size', 'dram__throughput.avg.pct']
for i, h in enumerate(H):
    if any(h.startswith(w) for w in want) and 'pcsamp' not in h and 'per_second' not in h and 'peak_sustained_elapsed' not in h.replace('sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed', '').replace('dram__throughput.avg.pct_of_peak_sustained_elapsed',''):
        vals = [r[i] for r in rows[2:]]
        if all(v in ('0', '', '0.000000') for v in vals): continue
        print(h, vals)
