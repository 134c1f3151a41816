#!/usr/bin/env python3
"""Per kernel of libminimod_cuda.so: SASS instruction count and the mnemonics that matter on this path.
usage: tools/sass_summary.py > profiles/rNN_sass_summary.txt   (cuobjdump -sass, no GPU needed)"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "minimod_b200", "lib", "libminimod_cuda.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True).stdout
WATCH = ("UBLKCP", "SYNCS", "REDG", "ATOMG", "LDG", "STG", "LDS", "STS", "SHFL", "VOTE", "POPC", "BAR", "LDL", "STL")
cur, stats = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); stats[cur] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line) if cur else None
    if m:
        stats[cur]["instr"] += 1
        base = m.group(2).split(".")[0]
        if base in WATCH:
            stats[cur][base] += 1
print("SASS of libminimod_cuda.so (cuobjdump -sass, sm_100a): instructions per kernel and selected mnemonics (static counts).")
print("UBLKCP = cp.async.bulk global->shared (bulk-copy engine); SYNCS = mbarrier arrive / try_wait; REDG = fire-and-forget reduction")
print("(the dense count cells); ATOMG = atomics with a return value (work counters, slot reservation); LDL/STL = local memory (spills).")
print("No tcgen05 / UTMALDG anywhere: nothing on this path is a contraction or a tiled tensor (DESIGN.md section 3).\n")
keys = ("instr",) + WATCH
print("%-58s " % "kernel" + " ".join("%6s" % k for k in keys))
for k, c in stats.items():
    name = subprocess.run(["c++filt", k], stdout=subprocess.PIPE, text=True).stdout.strip()
    name = re.sub(r"\(.*", "", name).replace("mmc::", "").replace("void ", "")
    print("%-58s " % name[:58] + " ".join("%6d" % c[x] for x in keys))
