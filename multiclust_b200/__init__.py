"""multiclust_b200 -- B200-native EM hot path of MULTICLUST.

The product is native: `libmc_cuda.so` (hand-written sm_100a kernels behind
the C ABI of include/mc_cuda.h) and the `host/multiclust` command line written
in C.  This package only carries a ctypes view of that ABI for the tests and
the benchmark; it contains no compute of its own and no CPU fallback.
"""
from .api import Context, McError, SynthParams, load_library, lib_path  # noqa: F401

__all__ = ["Context", "McError", "SynthParams", "load_library", "lib_path"]
