import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _make(*targets):
    subprocess.run(["make", "-C", ROOT, *targets], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)


@pytest.fixture(scope="session")
def host_lib():
    if not os.path.exists(os.path.join(ROOT, "minimod_b200", "lib", "libminimod_host.so")):
        _make("host")
    from minimod_b200 import _native
    return _native.load_host()


@pytest.fixture(scope="session")
def emul_lib(host_lib):
    """The kernel sources compiled for the CPU SIMT emulator (tests/kernel_emul): lets the CPU-only
    suite execute the real kernel code paths.  Test infrastructure, never a product fallback."""
    path = os.path.join(ROOT, "tests", "kernel_emul", "_build", "libminimod_emul.so")
    if not os.path.exists(path):
        _make("emul")
    from minimod_b200 import _native
    return _native.load_cuda(path)


@pytest.fixture(scope="session")
def cuda_lib(host_lib):
    import ctypes
    from minimod_b200 import _native
    lib = _native.load_cuda()          # raises if the .so is missing: no fallback
    return lib
