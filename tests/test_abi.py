"""The C-ABI library loads and exports every symbol include/minimod_cuda.h declares (no compute)."""
import ctypes
import os
import re

from helpers import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "minimod_cuda.h")).read()
    return sorted(set(re.findall(r"\b(mmc_[a-z0-9_]+)\s*\(", text)))


def test_header_and_bindings_agree():
    from minimod_b200 import _native
    assert declared_symbols() == sorted(_native.CUDA_SYMBOLS)


def test_cuda_library_exports_every_declared_symbol():
    from minimod_b200 import _native
    path = _native.cuda_lib_path()
    assert os.path.exists(path), "libminimod_cuda.so not built (make lib)"
    lib = ctypes.CDLL(path)
    for sym in declared_symbols():
        assert hasattr(lib, sym), sym
    assert lib.mmc_abi_version() == 2


def test_create_fails_loudly_without_a_device(host_lib):
    """On a CPU-only box the product library must refuse to run rather than fall back."""
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is present")
    import minimod_b200
    from helpers import DATA
    try:
        minimod_b200.Core("freq", os.path.join(DATA, "example-ont.bam"))
    except minimod_b200.MinimodError as e:
        assert "no CUDA device" in str(e) or "CUDA" in str(e)
    else:
        raise AssertionError("mmc_create succeeded without a CUDA device")
