"""Edge cases of the hot path on hand-built reads: the CUDA kernels (emulated on CPU here, on HBM under
-m gpu in test_gpu_edge_cases.py) must agree record-for-record with the oracle, and must flag exactly the
inputs the reference treats as fatal."""
import pytest

from edge_suite import CASES, FATAL, run_case, run_fatal


@pytest.mark.parametrize("case", CASES, ids=[c["id"] for c in CASES])
def test_edge_case_emulated(emul_lib, case):
    run_case(emul_lib, case)


@pytest.mark.parametrize("case", FATAL, ids=[c["id"] for c in FATAL])
def test_fatal_input_emulated(emul_lib, case):
    run_fatal(emul_lib, case)


def test_edge_cases_emulated_two_bit_seq(emul_lib, monkeypatch):
    """Every edge case again with SEQ in its 2-bit transport form (N and IUPAC bases, odd lengths -> exception
    entries); the oracle decodes the transport form with its own nibble-by-nibble loop."""
    monkeypatch.setenv("MMC_SEQ_PACKING", "2")
    for case in CASES:
        run_case(emul_lib, case)
    for case in FATAL:
        run_fatal(emul_lib, case)
