#!/usr/bin/env python3
"""Per source line (in line order): executed warp instructions and stall samples from an ncu report.
usage: ncu_lines_by_file.py report.ncu-rep [min_pct] [kernel-name-regex]"""
import csv, subprocess, sys
rep = sys.argv[1]; minp = float(sys.argv[2]) if len(sys.argv) > 2 else 0.2
kern = ["--kernel-name", "regex:" + sys.argv[3]] if len(sys.argv) > 3 else []
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"] + kern, stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = None; lines = {}; cur = None
for r in rows:
    if len(r) > 4 and r[0] == "Line No":
        hdr = r; ix_inst = hdr.index("Instructions Executed")
        ix_samp = hdr.index("# Samples") if "# Samples" in hdr else hdr.index("Warp Stall Sampling (All Samples)")
        continue
    if hdr is None or len(r) < len(hdr) - 5: continue
    if r[0].strip().isdigit():
        cur = (int(r[0]), r[1].strip()); lines.setdefault(cur, [0.0, 0.0]); continue
    if r[0] == "" and cur and r[2] not in ("...", "-"):
        try:
            lines[cur][0] += float(r[ix_inst] or 0); lines[cur][1] += float(r[ix_samp] or 0)
        except ValueError: pass
ti = sum(v[0] for v in lines.values()) or 1; ts = sum(v[1] for v in lines.values()) or 1
for (ln, src), (i, s) in sorted(lines.items()):
    if 100*i/ti >= minp or 100*s/ts >= minp:
        print(f"L{ln:<5d} {100*i/ti:5.1f}% inst {100*s/ts:5.1f}% samp  {src[:110]}")
