#!/usr/bin/env python3
"""Generate tests/golden/ from the reference checkout (run in the dev container only).

    python tests/golden/make_fixtures.py [/root/reference]

What it writes (all small, committed):
  data/*.bam            input fixtures copied from <ref>/test/data (binary test DATA, not source)
  expected/*.gz         the reference's own golden outputs from <ref>/test/expected
                        (commands: <ref>/test/test.sh:66-250), gzip-compressed
  pseudo_ref.npz        "poke lists" from which tests rebuild the pseudo reference genomes

Why a pseudo reference: the genomes the reference tests use are downloaded by wget
(<ref>/test/test.sh:31-37) and are not available offline.  For `[CG]`-context goldens the
only reference content that matters is where the CpGs are, and the golden *view* files list
exactly those positions:
  '+' row at pos  ->  C at pos,   G at pos+1
  '-' row at pos  ->  C at pos-1, G at pos
and for the single-base context `T` (eb.bam, -c e,b): '+' -> T at pos, '-' -> A at pos.
Everything else is 'N'.  With these genomes the unmodified reference binary
(oracle/_ref/minimod_ref) reproduces every golden file listed in GOLDEN_CASES
(tests/test_oracle_golden.py checks that on every run where the binary exists).
"""
import gzip
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

BAMS = [
    "example-ont.bam", "example-hifi.bam", "hap.bam", "eb.bam", "dRNA.bam",
    "dna_5mC_5hmC_mm_chr22.bam", "dna_6mA_mm_chr22.bam", "rna_algn_to_genome.bam",
]
EXPECTED = [
    "test1.tsv", "test2.tsv", "test2a.tsv", "test2b.tsv", "test2c.tsv", "test2c_wild.tsv",
    "test3.tsv", "test4.bedmethyl", "test5.tsv", "test5a.tsv", "test5b.tsv", "test5c.tsv",
    "test6.bedmethyl", "test7.tsv", "test8.tsv", "test9.tsv", "test10.tsv", "test11.tsv",
    "test12.tsv", "test15.tsv", "test16.tsv", "test17a.tsv",
]


def view_rows(path):
    with open(path) as fh:
        next(fh)
        for line in fh:
            f = line.split("\t")
            yield f[0], int(f[1]), f[2]


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    os.makedirs(os.path.join(HERE, "data"), exist_ok=True)
    os.makedirs(os.path.join(HERE, "expected"), exist_ok=True)
    for b in BAMS:
        shutil.copyfile(os.path.join(ref, "test/data", b), os.path.join(HERE, "data", b))
    for e in EXPECTED:
        with open(os.path.join(ref, "test/expected", e), "rb") as src:
            raw = src.read()
        with open(os.path.join(HERE, "expected", e + ".gz"), "wb") as raw_out:
            with gzip.GzipFile(filename="", mode="wb", fileobj=raw_out, mtime=0, compresslevel=9) as dst:
                dst.write(raw)

    pokes = {"chr22": {c: set() for c in "ACGT"}, "chr1": {c: set() for c in "ACGT"}}
    # CpG sites seen by `view -c m[CG]` (test1 hifi, test2 ont, test2c hap)
    for fn in ("test1.tsv", "test2.tsv", "test2c.tsv"):
        for contig, pos, strand in view_rows(os.path.join(ref, "test/expected", fn)):
            if strand == "+":
                pokes[contig]["C"].add(pos)
                pokes[contig]["G"].add(pos + 1)
            else:
                pokes[contig]["C"].add(pos - 1)
                pokes[contig]["G"].add(pos)
    # T sites seen by `view -c e,b` (context T)
    for contig, pos, strand in view_rows(os.path.join(ref, "test/expected", "test15.tsv")):
        pokes[contig]["T" if strand == "+" else "A"].add(pos)

    out = {}
    for contig, d in pokes.items():
        seen = {}
        for base, s in d.items():
            for p in s:
                assert seen.setdefault(p, base) == base, (contig, p, seen[p], base)
            out[f"{contig}_{base}"] = np.array(sorted(s), dtype=np.int32)
    np.savez_compressed(os.path.join(HERE, "pseudo_ref.npz"), **out)
    for k, v in out.items():
        print(k, len(v))


if __name__ == "__main__":
    main()
