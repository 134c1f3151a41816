"""Pin the oracle: the unmodified reference sources (oracle/_ref/minimod_ref, built by oracle/Makefile
against the htslib shim) must reproduce the reference's own golden files (test/test.sh:66-250) on the
pseudo reference genomes.  Raw-byte identity where rows are tie-free; identity after sorting otherwise
(the reference's tests sort both sides too)."""
import os

import pytest

from helpers import DATA, GOLDEN_CASES, golden_bytes, have_ref_bin, pseudo_fasta, run_ref, sorted_lines

pytestmark = pytest.mark.skipif(not have_ref_bin(), reason="oracle/_ref/minimod_ref not built (needs /root/reference)")

RAW_IDENTICAL = {"test1.tsv", "test2.tsv", "test2a.tsv", "test2b.tsv", "test2c_wild.tsv", "test2c.tsv", "test3.tsv",
                 "test4.bedmethyl", "test5.tsv", "test5a.tsv", "test5b.tsv", "test5c.tsv", "test6.bedmethyl", "test7.tsv",
                 "test8.tsv", "test9.tsv", "test10.tsv", "test11.tsv", "test12.tsv", "test15.tsv", "test17a.tsv"}


@pytest.mark.parametrize("name,sub,args,bam,contig", GOLDEN_CASES, ids=[c[0] for c in GOLDEN_CASES])
def test_reference_binary_reproduces_golden(name, sub, args, bam, contig):
    out = run_ref(sub, args, pseudo_fasta(contig), os.path.join(DATA, bam))
    gold = golden_bytes(name)
    if name in RAW_IDENTICAL:
        assert out == gold
    else:
        assert sorted_lines(out) == sorted_lines(gold)


def test_thread_count_does_not_change_output():
    fa, bam = pseudo_fasta("chr22"), os.path.join(DATA, "example-ont.bam")
    assert run_ref("freq", "-c m,h", fa, bam, threads=1) == run_ref("freq", "-c m,h", fa, bam, threads=8)
