"""CPU-only: emulated kernels vs oracle on small synthetic inputs of every BASELINE config shape."""
import pytest

from synth_suite import run_synth


@pytest.mark.parametrize("config,cov", [(2, 1.0), (3, 1.0), (6, 0.7), (4, 1.0)])
@pytest.mark.parametrize("sub", ["freq", "view"])
def test_synthetic_parity_emulated(emul_lib, config, cov, sub):
    n, st = run_synth(emul_lib, config, 120000, cov, sub)
    assert n > 0 and st.n_reads > 0
