"""mmc_freq_drain / mmc_freq_undrain (include/minimod_cuda.h): streaming read-back for coordinate-sorted input."""
import pytest

from drain_suite import check_drain_equals_finalize, check_drain_order_violation
from golden_runner import run_case
from helpers import GOLDEN_CASES, golden_bytes, sorted_lines


# ---- CPU: the SIMT-emulation build of the same sources
@pytest.mark.parametrize("config,lag", [(2, 1), (5, 2)])
def test_drains_plus_remainder_equal_one_finalize_emulated(emul_lib, config, lag):
    check_drain_equals_finalize(emul_lib, config, lag=lag, contig_len=60000, coverage=2.0)


@pytest.mark.parametrize("config", [3, 6])
def test_drain_with_side_buffer_records_emulated(emul_lib, config):
    """--insertions / '.' blocks with haplotypes: the side-buffer records below the watermark are sorted, reduced and merged
    into the drained rows while later batches may still be appending (their records read as sentinels until written)"""
    check_drain_equals_finalize(emul_lib, config, contig_len=40000, coverage=2.0)


def test_drain_order_violation_emulated(emul_lib):
    check_drain_order_violation(emul_lib, contig_len=60000, coverage=2.0)


def test_golden_through_drains_emulated(emul_lib):
    case = [c for c in GOLDEN_CASES if c[0] == "test7.tsv"][0]
    assert run_case(emul_lib, *case[1:], batch_size=16, drain=True) == golden_bytes("test7.tsv")


# ---- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("config,lag,chunks", [(2, 1, 6), (2, 2, 8), (4, 1, 5), (5, 2, 8)])
def test_drains_plus_remainder_equal_one_finalize(cuda_lib, config, lag, chunks):
    check_drain_equals_finalize(cuda_lib, config, lag=lag, chunks=chunks, contig_len=1500000, coverage=3.0 if config != 4 else 1.0)


@pytest.mark.gpu
@pytest.mark.parametrize("config,lag", [(3, 1), (3, 2), (6, 1)])
def test_drain_with_side_buffer_records(cuda_lib, config, lag):
    check_drain_equals_finalize(cuda_lib, config, lag=lag, chunks=8, contig_len=400000, coverage=2.0)


@pytest.mark.gpu
def test_drain_order_violation(cuda_lib):
    check_drain_order_violation(cuda_lib, contig_len=400000, coverage=2.0)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["test7.tsv", "test3.tsv", "test6.bedmethyl", "test5c.tsv", "test12.tsv", "test5a.tsv"])
def test_golden_through_drains(cuda_lib, name):
    case = [c for c in GOLDEN_CASES if c[0] == name][0]
    assert sorted_lines(run_case(cuda_lib, *case[1:], batch_size=8, drain=True)) == sorted_lines(golden_bytes(name))
