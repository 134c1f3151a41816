"""The drop-in proof: the reference's own drivers and loader (unmodified sources, compiled in place by oracle/Makefile) with
process_db / merge_db / output_db / output_core / load_ref_contexts supplied by integration/minimod_cuda_glue.c -- the
INTEGRATION.md binding made real -- must reproduce every golden file of the reference's test-suite
(/root/reference/test/test.sh:66-250).

  CPU (`not gpu`):  oracle/_ref/minimod_ref_emul, the binding linked to the SIMT-emulation build of the kernel sources
  GPU:              oracle/_ref/minimod_ref_cuda, the same objects linked to libminimod_cuda.so
"""
import os
import shlex
import subprocess

import pytest

from helpers import DATA, GOLDEN_CASES, ROOT, golden_bytes, pseudo_fasta, sorted_lines
from golden_runner import TIE_FREE

EMUL_BIN = os.path.join(ROOT, "oracle", "_ref", "minimod_ref_emul")
CUDA_BIN = os.path.join(ROOT, "oracle", "_ref", "minimod_ref_cuda")


def run_binding(binary, sub, args, bam, contig, extra=()):
    cmd = [binary, sub] + shlex.split(args) + list(extra) + ["-t", "4", pseudo_fasta(contig), os.path.join(DATA, bam)]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert res.returncode == 0, res.stderr.decode()[-1500:]
    return res.stdout


def check(binary, name, sub, args, bam, contig, extra=()):
    out, gold = run_binding(binary, sub, args, bam, contig, extra), golden_bytes(name)
    if name in TIE_FREE:
        assert out == gold, name
    else:
        assert sorted_lines(out) == sorted_lines(gold), name


CPU_SUBSET = {"test1.tsv", "test2a.tsv", "test2c.tsv", "test4.bedmethyl", "test5a.tsv", "test5c.tsv", "test7.tsv", "test16.tsv", "test17a.tsv"}


@pytest.mark.skipif(not os.path.exists(EMUL_BIN), reason="oracle/_ref/minimod_ref_emul not built (needs /root/reference)")
@pytest.mark.parametrize("name,sub,args,bam,contig", [c for c in GOLDEN_CASES if c[0] in CPU_SUBSET], ids=[c[0] for c in GOLDEN_CASES if c[0] in CPU_SUBSET])
def test_reference_with_glue_emulated(emul_lib, name, sub, args, bam, contig):
    """(a representative subset keeps the CPU suite short; the GPU suite runs all 22)"""
    check(EMUL_BIN, name, sub, args, bam, contig)


@pytest.mark.skipif(not os.path.exists(EMUL_BIN), reason="oracle/_ref/minimod_ref_emul not built (needs /root/reference)")
def test_reference_with_glue_small_batches(emul_lib):
    """-K 7: many batches in flight through the reference's load || process || merge pthread pipeline."""
    for name in ("test7.tsv", "test2.tsv", "test5c.tsv"):
        case = [c for c in GOLDEN_CASES if c[0] == name][0]
        check(EMUL_BIN, *case, extra=("-K", "7"))


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(CUDA_BIN), reason="oracle/_ref/minimod_ref_cuda not built (needs /root/reference)")
@pytest.mark.parametrize("name,sub,args,bam,contig", GOLDEN_CASES, ids=[c[0] for c in GOLDEN_CASES])
def test_reference_with_glue_cuda(name, sub, args, bam, contig):
    check(CUDA_BIN, name, sub, args, bam, contig)


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(CUDA_BIN), reason="oracle/_ref/minimod_ref_cuda not built (needs /root/reference)")
def test_reference_with_glue_cuda_small_batches():
    for name in ("test7.tsv", "test2.tsv", "test5c.tsv", "test5a.tsv"):
        case = [c for c in GOLDEN_CASES if c[0] == name][0]
        check(CUDA_BIN, *case, extra=("-K", "7"))
