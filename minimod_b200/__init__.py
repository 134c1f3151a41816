"""minimod_b200 -- B200-native decode + frequency aggregation hot path of warp9seq/minimod.

The product is libminimod_cuda.so (hand-written sm_100a kernels behind the C ABI of
include/minimod_cuda.h) plus the `minimod` command line (minimod_b200/bin/minimod).  This
package is the thin Python mirror of the reference's operator interface used by the tests and
bench.py; it holds no implementation of the algorithm itself.
"""
from .api import Core, freq, view, MinimodError  # noqa: F401

__version__ = "0.1.0"
