#!/usr/bin/env python3
"""One line per profiled launch of an ncu report: duration, registers, issue activity, resident warps, DRAM bytes, stalls.
usage: ncu_kernel_table.py report.ncu-rep"""
import csv, subprocess, sys
txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(txt.splitlines())); h, units = rows[0], rows[1]
def col(name): return h.index(name) if name in h else None
K = [("Kernel Name", 34), ("gpu__time_duration.sum", 9), ("launch__registers_per_thread", 5), ("smsp__issue_active.avg.pct_of_peak_sustained_active", 7),
     ("smsp__warps_active.avg.per_cycle_active", 6), ("dram__bytes_read.sum", 9), ("dram__bytes_write.sum", 9), ("smsp__inst_executed.sum", 12)]
print("kernel                              dur(%s)  regs  issue%%  warps/sched  dram_rd  dram_wr  warp_instr   top stalls (warps per issue)" % units[col("gpu__time_duration.sum")])
for r in rows[2:]:
    out = []
    for name, w in K:
        i = col(name); v = r[i] if i is not None else "-"
        if name == "Kernel Name": v = v.split("(")[0][:w]
        elif name.startswith("dram"): v = "%.3f%s" % (float(v), units[i][:1]) if v not in ("", "-") else v
        else:
            try: v = "%.1f" % float(v) if "." in v else v
            except ValueError: pass
        out.append(v.ljust(w))
    st = {k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): float(r[i]) for i, k in enumerate(h)
          if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and r[i] not in ("", "n/a")}
    top = ", ".join(f"{k} {x:.2f}" for k, x in sorted(st.items(), key=lambda kv: -kv[1])[:4])
    print("  ".join(out) + "  " + top)
