"""pytest configuration: markers, import path, shared fixtures."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def orc():
    """the CPU oracle (test infrastructure; oracle/mc_oracle.h)"""
    from oracle import orc as o
    o.lib()
    return o


@pytest.fixture(scope="session")
def mclib():
    """libmc_cuda.so built in-tree; never falls back to anything else"""
    from multiclust_b200 import build, api
    if not os.path.exists(api.lib_path()):
        build.build_cuda()
    return api.load_library()
