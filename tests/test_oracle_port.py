"""Pin the C restatement (oracle/modcall_oracle.c): it must reproduce every golden file of the
reference's own tests that covers this path, and agree with the unmodified reference binary on
fixtures the goldens do not cover ('.' status blocks, insertions on them, 6mA, wildcard code)."""
import os
import shlex

import pytest

import oracle_port
from helpers import DATA, GOLDEN_CASES, golden_bytes, have_ref_bin, pseudo_fasta, run_ref, sorted_lines


def run_oracle(sub, args, bam, contig):
    kw, codes, thresh = {}, "m", None
    toks = shlex.split(args)
    i = 0
    while i < len(toks):
        t = toks[i]
        if t == "-c":
            codes = toks[i + 1]; i += 1
        elif t == "-m":
            thresh = toks[i + 1]; i += 1
        elif t == "-b":
            kw["bedmethyl"] = True
        elif t == "-K":
            kw["batch_size"] = int(toks[i + 1]); i += 1
        elif t == "--insertions":
            kw["insertions"] = True
        elif t == "--haplotypes":
            kw["haplotypes"] = True
        i += 1
    return oracle_port.run(sub, pseudo_fasta(contig), os.path.join(DATA, bam), codes, thresh, **kw)


@pytest.mark.parametrize("name,sub,args,bam,contig", GOLDEN_CASES, ids=[c[0] for c in GOLDEN_CASES])
def test_restatement_reproduces_golden(name, sub, args, bam, contig):
    assert sorted_lines(run_oracle(sub, args, bam, contig)) == sorted_lines(golden_bytes(name))


@pytest.mark.skipif(not have_ref_bin(), reason="oracle/_ref/minimod_ref not present")
@pytest.mark.parametrize("bam,args", [
    ("dna_5mC_5hmC_mm_chr22.bam", "-c m[*],h[*]"),
    ("dna_5mC_5hmC_mm_chr22.bam", "-c m[*],h[*] --insertions"),
    ("dna_6mA_mm_chr22.bam", "-c a[*]"),
    ("dna_6mA_mm_chr22.bam", "-c *"),
    ("rna_algn_to_genome.bam", "-c a[*]"),
])
@pytest.mark.parametrize("sub", ["freq", "view"])
def test_restatement_matches_reference_binary(bam, args, sub):
    fa, path = pseudo_fasta("chr22"), os.path.join(DATA, bam)
    assert sorted_lines(run_oracle(sub, args, bam, "chr22")) == sorted_lines(run_ref(sub, args, fa, path))
