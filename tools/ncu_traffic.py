#!/usr/bin/env python3
"""profiles/traffic.json from ncu captures of the decode stage: dram__bytes_read.sum + dram__bytes_write.sum of the
stage's kernels (k_flat_setup + the decode kernel), per launch of the whole shard.  bench.py reports it as roofline.traffic.
usage: ncu_traffic.py config:report.ncu-rep:reads ..."""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path = os.environ.get("TRAFFIC_JSON", os.path.join(ROOT, "profiles", "traffic.json"))
try:
    out = json.load(open(path))
except Exception:
    out = {}
for spec in sys.argv[1:]:
    cfg, rep, reads = spec.split(":")
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(txt.splitlines())); h, u = rows[0], rows[1]
    def val(r, k):
        i = h.index(k); x = float(r[i]); unit = u[i]
        return x * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[unit]
    seen, total, parts = set(), 0.0, {}
    for r in rows[2:]:
        name = r[h.index("Kernel Name")].split("(")[0].replace("void ", "")
        if name in seen:                       # first launch of each kernel of the stage
            continue
        seen.add(name)
        b = val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
        parts[name] = int(b); total += b
    out[f"config{cfg}"] = {"dram_bytes_per_launch": int(total), "reads": int(reads), "kernels": parts,
                          "source": "ncu --set full --clock-control none, " + os.path.basename(rep)}
    print(cfg, int(total), parts)
json.dump(out, open(path, "w"), indent=1, sort_keys=True)
