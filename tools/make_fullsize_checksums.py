#!/usr/bin/env python3
"""Digest of the freq table the UNMODIFIED reference (oracle/_ref/minimod_ref) prints for the full-size synthetic jobs of
BASELINE.json configs 2-5 -> tests/golden/fullsize_checksums.json.  Run where /root/reference exists (this container);
the GPU box only has the committed JSON.  usage: tools/make_fullsize_checksums.py [config ...]"""
import json, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import fullsize
from helpers import REF_BIN

todo = [int(a) for a in sys.argv[1:]] or sorted(fullsize.JOBS)
try:
    done = {str(k): v for k, v in fullsize.load_checksums().items()}
except Exception:
    done = {}
for c in todo:
    with tempfile.TemporaryDirectory(prefix="mm_full_", dir=os.environ.get("MM_TMP", "/tmp")) as td:
        t0 = time.time()
        fa, bam, args, st = fullsize.write_job(c, td)
        t1 = time.time()
        dg, err = fullsize.digest_of_command([REF_BIN, "freq"] + args + ["-t", str(os.cpu_count()), "-K", "4092", "-B", "100M", fa, bam])
        dg.update(reads=int(st["n_reads"]), bases=int(st["bases"]), coverage=fullsize.JOBS[c], args=" ".join(args),
                  reference_seconds=round(time.time() - t1, 1), reference="oracle/_ref/minimod_ref (unmodified /root/reference sources) -t %d" % os.cpu_count())
        done[str(c)] = dg
        print(c, dg, "gen %.0f s" % (t1 - t0), flush=True)
        with open(fullsize.CHECKSUMS, "w") as fh:
            json.dump(done, fh, indent=1, sort_keys=True)
