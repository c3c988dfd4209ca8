import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")

# The xBLAT3 testers replace XERBLA (blas/testing/dblat3.f:2818-2850).  ctypes binds symbols at load time
# (RTLD_NOW), so the tester's xerbla_ (oracle/blat3_port.c, exported RTLD_GLOBAL) has to be in the global scope
# BEFORE libb200blas.so / the reference library are loaded; unarmed it prints exactly like blas/xerbla.cpp.
import oracle_api  # noqa: E402

oracle_api.port()
