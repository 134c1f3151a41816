#!/usr/bin/env python3
"""Summarise an ncu report's source page per CUDA source line: share of executed warp instructions
and of stall samples.  usage: ncu_hot_lines.py report.ncu-rep [top_n]"""
import csv
import subprocess
import sys

rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = None
lines = {}
cur = None
for r in rows:
    if len(r) > 4 and r[0] == "Line No":
        hdr = r
        ix_inst = hdr.index("Instructions Executed")
        ix_samp = hdr.index("# Samples") if "# Samples" in hdr else hdr.index("Warp Stall Sampling (All Samples)")
        continue
    if hdr is None or len(r) < len(hdr) - 5:
        continue
    if r[0].strip().isdigit():
        cur = (int(r[0]), r[1].strip())
        lines.setdefault(cur, [0.0, 0.0])
        continue
    if r[0] == "" and cur and r[2] not in ("...", "-"):
        try:
            lines[cur][0] += float(r[ix_inst] or 0)
            lines[cur][1] += float(r[ix_samp] or 0)
        except ValueError:
            pass
ti = sum(v[0] for v in lines.values()) or 1
ts = sum(v[1] for v in lines.values()) or 1
print(f"total warp instructions {ti:.3g}, stall samples {ts:.0f}")
print("== by executed instructions")
for (ln, src), (i, s) in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100 * i / ti:5.1f}% inst {100 * s / ts:5.1f}% samp  L{ln:<5d} {src[:100]}")
print("== by stall samples")
for (ln, src), (i, s) in sorted(lines.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{100 * i / ti:5.1f}% inst {100 * s / ts:5.1f}% samp  L{ln:<5d} {src[:100]}")
