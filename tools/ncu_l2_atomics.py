#!/usr/bin/env python3
"""L2 reduction / atomic counters of every profiled launch in an ncu report (the N1 question: would privatised
shared-memory histograms help?).  usage: ncu_l2_atomics.py report.ncu-rep"""
import csv, subprocess, sys
txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(txt.splitlines())); h, u = rows[0], rows[1]
W = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
     "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum", "lts__t_requests_srcunit_tex_op_red.sum",
     "lts__t_sectors_srcunit_tex_op_red.sum", "lts__t_sectors_srcunit_tex_op_red_lookup_hit.sum", "lts__t_sectors_srcunit_tex_op_red_lookup_miss.sum",
     "lts__t_sectors_srcunit_tex_op_red.sum.pct_of_peak_sustained_elapsed", "lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed",
     "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors.sum"]
for r in rows[2:]:
    print(r[h.index("Kernel Name")].split("(")[0])
    for k in W:
        if k in h:
            print("    %-82s %14s %s" % (k, r[h.index(k)], u[h.index(k)]))
