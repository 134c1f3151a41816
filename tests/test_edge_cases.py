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


def test_edge_cases_emulated_byte_cigar(emul_lib, monkeypatch):
    """Every edge case again with the CIGARs in their byte form (lengths 0..14 in the nibble, 15..269 in the first escape
    list, the huge ops of `huge_cigar_ops` in the second); the oracle decodes the form with its own sequential loop."""
    monkeypatch.setenv("MMC_CIGAR_PACKING", "8")
    monkeypatch.setenv("MMC_SEQ_PACKING", "2")
    for case in CASES:
        run_case(emul_lib, case)
    for case in FATAL:
        run_fatal(emul_lib, case)


def test_view_buffer_regrows(emul_lib):
    """A view batch with more rows than the slot's record buffer: the buffer is regrown and the batch re-run
    (it used to fail with 'raise view_capacity', an option the CLI does not have)."""
    import ctypes as C
    from minimod_b200.synth import Synth
    from parity import Pair
    s = Synth(6, contigs=(("chrS", 60000),), coverage=1.0)
    p, n = s.ref(0)
    pair = Pair(emul_lib, "view", [("chrS", C.string_at(p, n))], "m[*],h[*]", None, max_reads=s.n_reads + 8, max_bytes=8 << 20, view_capacity=7)
    try:
        got, _ = s.fill(pair.batch, 0, s.n_reads, 2)
        assert got == s.n_reads
        rc, msg = pair.run_device(); assert rc == 0, msg
        orc, omsg = pair.run_oracle(); assert orc == 0, omsg
        d, o = pair.device_view(), pair.oracle_view()
        assert len(d) > 100 and d == o
    finally:
        pair.close(); s.close()
