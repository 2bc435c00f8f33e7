"""The C-ABI library loads on a CPU-only box and exports every symbol that
include/mc_cuda.h declares (no compute calls here)."""
import ctypes
import os
import re

from multiclust_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = []
    for hdr in ("mc_cuda.h",):
        text = open(os.path.join(ROOT, "include", hdr)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names += re.findall(r"\b(mc_[A-Za-z0-9_]+)\s*\(", text)
    return sorted(set(names))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(api.SYMBOLS)


def test_library_exports_every_symbol(mclib):
    for name in declared_symbols():
        assert hasattr(mclib, name), name
    assert mclib.mc_abi_version() == 3


def test_no_cpu_fallback_without_device(mclib):
    """mc_create must fail loudly when there is no CUDA device"""
    import torch
    if torch.cuda.is_available():
        return
    h = ctypes.c_void_p()
    rc = mclib.mc_create(ctypes.byref(h), 0)
    assert rc != 0
    assert b"no CPU fallback" in mclib.mc_last_error(None)
