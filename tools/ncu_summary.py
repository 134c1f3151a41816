#!/usr/bin/env python3
"""Summarise an ncu report: opcode mix, warp instructions per read, stall reasons, hot source lines.
usage: ncu_summary.py report.ncu-rep [reads_per_launch] [min_line_pct]"""
import csv, collections, subprocess, sys
rep = sys.argv[1]; reads = float(sys.argv[2]) if len(sys.argv) > 2 else 99616.0
def page(args):
    return list(csv.reader(subprocess.run(["ncu", "-i", rep] + args + ["--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout.splitlines()))
rows = page(["--page", "source", "--print-source", "sass"])
hdr = rows[1]; ii = hdr.index("Instructions Executed"); it = hdr.index("Thread Instructions Executed")
tot = tt = 0; ops = collections.Counter(); static = 0
for r in rows[2:]:
    if len(r) <= ii: continue
    try: n = float(r[ii]); t = float(r[it])
    except ValueError: continue
    static += 1; tot += n; tt += t
    toks = r[1].split(); op = (toks[1] if toks[0].startswith('@') else toks[0]).split('.')[0]; ops[op] += n
print(f"static SASS {static}; warp instr {tot:.4g} = {tot/reads:.0f}/read; avg active lanes {tt/max(tot,1):.1f}")
print("opcodes: " + ", ".join(f"{o} {100*n/tot:.1f}%" for o, n in ops.most_common(16)))
raw = page(["--page", "raw"])
h = raw[0]; v = raw[2] if len(raw) > 2 else raw[1]
d = dict(zip(h, v))
keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum"]
for k in keys:
    if k in d: print(f"  {k} = {d[k]}")
st = {k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): float(x) for k, x in d.items()
      if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")}
print("stalls (warps per issue): " + ", ".join(f"{k} {x:.2f}" for k, x in sorted(st.items(), key=lambda kv: -kv[1])[:8]))
