"""Run one reference test.sh case through the C ABI (minimod_b200.freq / .view) with a given library."""
import os
import shlex

import minimod_b200
from helpers import DATA, pseudo_fasta

# rows that share (contig,pos) come out of the reference in hash-table order: the
# reference's tests sort before diffing, so do we -- but single-code, context-checked outputs are
# tie-free and must be raw-byte identical.
TIE_FREE = {"test1.tsv", "test2.tsv", "test2a.tsv", "test2b.tsv", "test2c.tsv", "test3.tsv", "test4.bedmethyl",
            "test5.tsv", "test6.bedmethyl", "test7.tsv", "test9.tsv", "test10.tsv", "test17a.tsv"}


def run_case(lib, sub, args, bam, contig, **extra):
    kw = dict(extra)
    toks = shlex.split(args)
    codes, thresh, bed = "m", None, False
    i = 0
    while i < len(toks):
        t = toks[i]
        if t == "-c":
            codes = toks[i + 1]; i += 1
        elif t == "-m":
            thresh = toks[i + 1]; i += 1
        elif t == "-b":
            bed = True
        elif t == "-K":
            kw["batch_size"] = int(toks[i + 1]); i += 1
        elif t == "--insertions":
            kw["insertions"] = True
        elif t == "--haplotypes":
            kw["haplotypes"] = True
        else:
            raise ValueError(t)
        i += 1
    fa, bam_path = pseudo_fasta(contig), os.path.join(DATA, bam)
    if sub == "freq":
        return minimod_b200.freq(fa, bam_path, codes, thresh, bedmethyl=bed, lib=lib, **kw)
    return minimod_b200.view(fa, bam_path, codes, lib=lib, **kw)
