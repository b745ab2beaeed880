#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --page raw --csv) into the handful of numbers we track."""
import csv, subprocess, sys, io
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_tex_lookup_hit.sum', 'lts__t_sectors_srcunit_tex_lookup_miss.sum']
for r in data:
    print("-" * 60)
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print("%-70s %s %s" % (w, r[i][:90], units[i]))
    for i, h in enumerate(hdr):
        if 'average_warps_issue_stalled' in h and h.endswith('per_issue_active.ratio'):
            try:
                v = float(r[i])
            except ValueError:
                continue
            if v > 0.5:
                print("%-70s %.2f" % (h.replace('smsp__average_warps_issue_stalled_', 'stall:').replace('_per_issue_active.ratio', ''), v))
