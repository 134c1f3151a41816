#!/usr/bin/env python3
"""Per-source-line instruction counts and stall samples of one kernel out of an ncu report
(`ncu -i X.ncu-rep --page source --csv --print-source cuda,sass -k regex:NAME`), as a table
sorted by executed warp instructions.  Usage: ncu_lines.py file.csv [kernel-substring] [top]"""
import csv, sys, collections
path = sys.argv[1]; want = sys.argv[2] if len(sys.argv) > 2 else ""; top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
rows = list(csv.reader(open(path, newline="")))
kern = None; fname = None; hdr = None
agg = collections.OrderedDict()
for r in rows:
    if not r: continue
    if r[0] in ("Kernel Name", "Function Name"): kern = r[1]; hdr = None; continue
    if r[0] in ("File Name", "File Path"): fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or want not in (kern or ""): continue
    if r[2] != "-": continue            # SASS rows carry an address; line rows have "-"
    d = dict(zip(hdr[4:], r[4:]))
    try:
        inst = int(d["Instructions Executed"]); samp = int(d["# Samples"])
    except Exception: continue
    key = (kern, fname, int(r[0]))
    a = agg.setdefault(key, [0, 0, r[1], 0, 0])
    a[0] += inst; a[1] += samp
    a[3] += int(d.get("stall_long_sb", 0) or 0); a[4] += int(d.get("stall_wait", 0) or 0)
bykern = collections.defaultdict(list)
for (k, f, l), a in agg.items(): bykern[k].append((a[0], a[1], f, l, a[2], a[3], a[4]))
for k, v in bykern.items():
    ti = sum(x[0] for x in v); ts = sum(x[1] for x in v)
    print(f"== {k}\n   warp instructions {ti}, samples {ts}")
    print("   inst%  samp%  long_sb  wait   file:line  source")
    for inst, samp, f, l, src, lsb, wt in sorted(v, reverse=True)[:top]:
        print(f"   {100*inst/max(ti,1):5.2f}  {100*samp/max(ts,1):5.2f}  {lsb:6d} {wt:6d}  {f}:{l}  {src.strip()[:110]}")
