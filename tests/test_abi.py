"""CPU-only checks of the drop-in boundary: libb200blas.so builds, loads, exports every symbol that
include/b200blas.h declares, and takes the reference's error exits / quick returns before touching CUDA
(blas/level3_impl.h:47-60).  No compute calls here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import eigen_b200
import oracle_api as oa

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported():
    hdr = open(os.path.join(ROOT, "include", "b200blas.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b([a-z_0-9]+_?)\s*\(", hdr)) - {"defined"}
    declared = {d for d in declared if d.endswith("gemm_") or d.startswith("b200blas_") or d == "xerbla_"}
    assert declared == set(eigen_b200.EXPORTS), declared ^ set(eigen_b200.EXPORTS)
    L = eigen_b200.lib()
    for name in declared:
        assert hasattr(L, name), name


def test_version_and_no_device_is_loud():
    L = eigen_b200.lib()
    assert L.b200blas_version() >= 100
    import torch
    if not torch.cuda.is_available():
        assert L.b200blas_device_ok() == 0
        with pytest.raises(RuntimeError):
            eigen_b200.require_device()


@pytest.mark.parametrize("t", list("sdcz"))
def test_error_exits_without_gpu(t):
    """xCHKE's 28 illegal GEMM calls (dblat3.f:1889-1972): every one must reach xerbla_ with the right info and
    name and leave C alone.  Argument checking precedes any CUDA work, so this runs on a CPU-only box."""
    P = oa.port()  # RTLD_GLOBAL: its xerbla_ (the tester's) interposes the library's weak one
    L = C.CDLL(eigen_b200.LIB_PATH, mode=C.RTLD_LOCAL)
    fn = getattr(L, t + "gemm_")
    log = C.create_string_buffer(4096)
    failed = P.oracle_blat3_chke(oa.TYPES[t], C.cast(fn, C.c_void_p), None, log, 4096)
    assert failed == 0, log.value.decode()


@pytest.mark.parametrize("t", list("sdcz"))
def test_quick_return_m0_n0(t):
    """m == 0 or n == 0 returns before anything is touched (blas/level3_impl.h:59-60) -- also without a GPU."""
    L = eigen_b200.lib()
    dt = oa.NP_DTYPE[t]
    a = np.ones((4, 4), dtype=dt, order="F")
    c = np.full((4, 4), 7, dtype=dt, order="F")
    for (m, n) in ((0, 3), (3, 0), (0, 0)):
        r = oa.call_gemm(getattr(L, t + "gemm_"), t, "N", "N", m, n, 2, 1.0, a, 4, a, 4, 0.0, c, 4)
        assert r == 0
        assert np.all(c == 7)
