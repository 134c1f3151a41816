import pytest

from edge_suite import CASES, FATAL, run_case, run_fatal

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", CASES, ids=[c["id"] for c in CASES])
def test_edge_case_cuda(cuda_lib, case):
    run_case(cuda_lib, case)


@pytest.mark.parametrize("case", FATAL, ids=[c["id"] for c in FATAL])
def test_fatal_input_cuda(cuda_lib, case):
    run_fatal(cuda_lib, case)


def test_edge_cases_cuda_two_bit_seq(cuda_lib, monkeypatch):
    """Every edge case again with SEQ in its 2-bit transport form (N and IUPAC bases, odd lengths -> exception
    entries; k_unpack_seq2 / k_patch_seq4 rebuild the 4-bit pool on the device)."""
    monkeypatch.setenv("MMC_SEQ_PACKING", "2")
    for case in CASES:
        run_case(cuda_lib, case)
    for case in FATAL:
        run_fatal(cuda_lib, case)


def test_edge_cases_cuda_byte_cigar(cuda_lib, monkeypatch):
    """Every edge case again with the CIGARs in their byte form (k_unpack_cigar rebuilds the word pool on the device)."""
    monkeypatch.setenv("MMC_CIGAR_PACKING", "8")
    monkeypatch.setenv("MMC_SEQ_PACKING", "2")
    for case in CASES:
        run_case(cuda_lib, case)
    for case in FATAL:
        run_fatal(cuda_lib, case)
