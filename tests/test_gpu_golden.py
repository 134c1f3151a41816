"""GPU parity, through the C ABI (libminimod_cuda.so loaded in this process) and through the `minimod`
binary: every golden file of the reference's test.sh that pins this path."""
import os
import shlex
import subprocess

import pytest

from golden_runner import TIE_FREE, run_case
from helpers import DATA, GOLDEN_CASES, ROOT, golden_bytes, have_ref_bin, pseudo_fasta, run_ref, sorted_lines

pytestmark = pytest.mark.gpu
CLI = os.path.join(ROOT, "minimod_b200", "bin", "minimod")


@pytest.mark.parametrize("name,sub,args,bam,contig", GOLDEN_CASES, ids=[c[0] for c in GOLDEN_CASES])
def test_cuda_reproduces_golden(cuda_lib, name, sub, args, bam, contig):
    out = run_case(cuda_lib, sub, args, bam, contig)
    gold = golden_bytes(name)
    if name in TIE_FREE:
        assert out == gold
    else:
        assert sorted_lines(out) == sorted_lines(gold)


@pytest.mark.parametrize("name", ["test7.tsv", "test5a.tsv", "test2a.tsv", "test17a.tsv", "test5c.tsv"])
def test_cuda_scratch_paths(cuda_lib, monkeypatch, name):
    monkeypatch.setenv("MMC_TEST_SMALL_SMEM", "1")
    case = [c for c in GOLDEN_CASES if c[0] == name][0]
    out = run_case(cuda_lib, *case[1:])
    assert sorted_lines(out) == sorted_lines(golden_bytes(name))


@pytest.mark.parametrize("threads", ["32", "64", "256"])
def test_cuda_cta_sizes(cuda_lib, monkeypatch, threads):
    monkeypatch.setenv("MMC_DECODE_THREADS", threads)
    for name in ("test7.tsv", "test8.tsv"):
        case = [c for c in GOLDEN_CASES if c[0] == name][0]
        assert sorted_lines(run_case(cuda_lib, *case[1:])) == sorted_lines(golden_bytes(name))


@pytest.mark.parametrize("name", ["test7.tsv", "test4.bedmethyl", "test2.tsv", "test12.tsv"])
def test_cli_binary(name):
    case = [c for c in GOLDEN_CASES if c[0] == name][0]
    _, sub, args, bam, contig = case
    res = subprocess.run([CLI, sub] + shlex.split(args) + [pseudo_fasta(contig), os.path.join(DATA, bam)],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert res.returncode == 0, res.stderr.decode()[-2000:]
    if name in TIE_FREE:
        assert res.stdout == golden_bytes(name)
    else:
        assert sorted_lines(res.stdout) == sorted_lines(golden_bytes(name))


@pytest.mark.skipif(not have_ref_bin(), reason="oracle/_ref/minimod_ref not present")
@pytest.mark.parametrize("bam,args", [
    ("dna_5mC_5hmC_mm_chr22.bam", "-c m[*],h[*]"),          # '.' status blocks: implicit calls
    ("dna_5mC_5hmC_mm_chr22.bam", "-c m[*],h[*] --insertions"),
    ("dna_6mA_mm_chr22.bam", "-c a[*]"),
    ("dna_6mA_mm_chr22.bam", "-c *"),
])
def test_cuda_vs_reference_binary_live(cuda_lib, bam, args):
    fa, path = pseudo_fasta("chr22"), os.path.join(DATA, bam)
    for sub in ("freq", "view"):
        ref_out = run_ref(sub, args, fa, path)
        out = run_case(cuda_lib, sub, args, bam, "chr22")
        assert sorted_lines(out) == sorted_lines(ref_out)


PATHS = [("stream", None, "3"), ("split", None, "3"), ("split", None, "4"), ("split", None, "2"), ("warp", None, "3"), ("general", None, "3"),
         ("split", "4608", "3"), ("warp", "4400", "4")]


@pytest.mark.parametrize("path,arena,occ", PATHS, ids=[f"{p}-{a or 'default'}-occ{o}" for p, a, o in PATHS])
def test_cuda_decode_paths_agree(cuda_lib, monkeypatch, path, arena, occ):
    """Streaming, split, warp-per-read and general kernels, arena sizes that force sampling / deferral."""
    monkeypatch.setenv("MMC_DECODE_PATH", path)
    monkeypatch.setenv("MMC_WARP_OCC", occ)
    if arena:
        monkeypatch.setenv("MMC_WARP_ARENA", arena)
    for name in ("test7.tsv", "test5a.tsv", "test5c.tsv", "test17a.tsv", "test16.tsv", "test2b.tsv", "test11.tsv"):
        case = [c for c in GOLDEN_CASES if c[0] == name][0]
        assert sorted_lines(run_case(cuda_lib, *case[1:])) == sorted_lines(golden_bytes(name)), name


@pytest.mark.gpu
def test_cuda_sparse_rows_sorted_on_device(cuda_lib, monkeypatch):
    """Every freq golden with the sparse side buffer sorted, reduced and merged on the device (cub radix sort +
    mmc_sparse.cuh) whatever its size; the default takes that path only above 65536 records."""
    monkeypatch.setenv("MMC_SPARSE_DEVICE_MIN", "0")
    for name in ("test17a.tsv", "test5a.tsv", "test5c.tsv", "test7.tsv", "test16.tsv", "test2b.tsv", "test11.tsv"):
        case = [c for c in GOLDEN_CASES if c[0] == name][0]
        assert sorted_lines(run_case(cuda_lib, *case[1:])) == sorted_lines(golden_bytes(name)), name


@pytest.mark.gpu
def test_cuda_goldens_four_bit_seq(cuda_lib, monkeypatch):
    """The Python mirror sends SEQ at 2 bits per base by default; the freq goldens once more with BAM's 4-bit nibbles."""
    monkeypatch.setenv("MMC_SEQ_PACKING", "4")
    for name in ("test17a.tsv", "test5a.tsv", "test5c.tsv", "test7.tsv", "test16.tsv", "test2b.tsv", "test11.tsv"):
        case = [c for c in GOLDEN_CASES if c[0] == name][0]
        assert sorted_lines(run_case(cuda_lib, *case[1:])) == sorted_lines(golden_bytes(name)), name
