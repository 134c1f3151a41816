"""CPU-only: the product's kernel sources, compiled for the SIMT emulator in tests/kernel_emul, driven
through the same C ABI + host code as on the GPU, against the reference's golden files.  This is what
lets `pytest -m "not gpu"` exercise the decode/finalize kernels' logic; the GPU tests repeat it on HBM."""
import pytest

from golden_runner import TIE_FREE, run_case
from helpers import GOLDEN_CASES, golden_bytes, sorted_lines

CHR22 = [c for c in GOLDEN_CASES if c[4] == "chr22"]
CHR1 = [c for c in GOLDEN_CASES if c[4] == "chr1"]


@pytest.mark.parametrize("name,sub,args,bam,contig", CHR22 + CHR1, ids=[c[0] for c in CHR22 + CHR1])
def test_emulated_kernels_reproduce_golden(emul_lib, name, sub, args, bam, contig):
    out = run_case(emul_lib, sub, args, bam, contig)
    gold = golden_bytes(name)
    if name in TIE_FREE:
        assert out == gold
    else:
        assert sorted_lines(out) == sorted_lines(gold)


@pytest.mark.parametrize("name", ["test7.tsv", "test5a.tsv", "test2a.tsv", "test17a.tsv"])
def test_scratch_paths(emul_lib, monkeypatch, name):
    """Force the global-scratch fallbacks (long CIGARs / long reads) with tiny shared-memory caps."""
    monkeypatch.setenv("MMC_TEST_SMALL_SMEM", "1")
    case = [c for c in GOLDEN_CASES if c[0] == name][0]
    out = run_case(emul_lib, *case[1:])
    assert sorted_lines(out) == sorted_lines(golden_bytes(name))


# The decode stage has three implementations that must agree: the split path (k_flat_setup + k_decode_warp<PRE>,
# the default), the self-contained warp-per-read kernel, and the general CTA-per-read kernel (what the
# other two defer reads to).  MMC_WARP_ARENA shrinks the per-warp shared-memory arena so that CIGAR / index
# sampling and deferral actually happen on the small fixtures.
PATHS = [("warp", None), ("general", None), ("split", "4608"), ("warp", "4400"), ("stream", None), ("split", None)]   # (default: per batch, by read shape)


@pytest.mark.parametrize("path,arena", PATHS, ids=[f"{p}-{a or 'default'}" for p, a in PATHS])
@pytest.mark.parametrize("name", ["test7.tsv", "test5a.tsv", "test17a.tsv", "test16.tsv", "test5c.tsv"])
def test_decode_paths_agree(emul_lib, monkeypatch, path, arena, name):
    if path != "stream" and name in ("test16.tsv", "test5c.tsv"):
        pytest.skip("the extra cases are for the streaming kernel; the GPU suite runs every path on every case")
    monkeypatch.setenv("MMC_DECODE_PATH", path)
    if arena:
        monkeypatch.setenv("MMC_WARP_ARENA", arena)
    case = [c for c in GOLDEN_CASES if c[0] == name][0]
    out = run_case(emul_lib, *case[1:])
    assert sorted_lines(out) == sorted_lines(golden_bytes(name))


@pytest.mark.parametrize("name", ["test17a.tsv", "test5a.tsv"])
def test_sparse_rows_sorted_on_device(emul_lib, monkeypatch, name):
    """--insertions / haplotype rows through the device-side sort + merge of the sparse side buffer
    (mmc_sparse.cuh; normally only taken above 65536 records) instead of the host sort."""
    monkeypatch.setenv("MMC_SPARSE_DEVICE_MIN", "0")
    case = [c for c in GOLDEN_CASES if c[0] == name][0]
    out = run_case(emul_lib, *case[1:])
    assert sorted_lines(out) == sorted_lines(golden_bytes(name))


@pytest.mark.parametrize("name", ["test8.tsv", "test12.tsv", "test11.tsv", "test5a.tsv", "test16.tsv", "test2c_wild.tsv"])
def test_split_blocks_agree(emul_lib, monkeypatch, name):
    """k_decode_stream, split mode: the even and the odd MM blocks of a read are decoded by different warps, each block's
    first ML index (src/mod.c:1200) precomputed by k_flat_setup from the token counts of the blocks before it."""
    monkeypatch.setenv("MMC_DECODE_PATH", "stream")
    monkeypatch.setenv("MMC_STREAM_SPLIT", "1")
    case = [c for c in GOLDEN_CASES if c[0] == name][0]
    out = run_case(emul_lib, *case[1:])
    assert sorted_lines(out) == sorted_lines(golden_bytes(name))
