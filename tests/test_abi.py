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
    blas3 = ("gemm_", "syrk_", "herk_", "trsm_", "trmm_", "symm_", "hemm_", "syr2k_", "her2k_", "potrf_", "getrf_")
    declared = {d for d in declared if (d.endswith(blas3) and len(d) <= 7) or d.startswith("b200blas_") or d == "xerbla_"}
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


@pytest.mark.parametrize("name", oa.RANKK_NAMES)
def test_rankk_error_exits_without_gpu(name):
    """?SYRK / ?HERK blocks of xCHKE (info 1, 2, 3, 4, 7, 10) -- argument checking precedes any CUDA work."""
    P = oa.port()
    L = C.CDLL(eigen_b200.LIB_PATH, mode=C.RTLD_LOCAL)
    for nm in oa.RANKK_NAMES:
        getattr(L, nm).argtypes = oa.RANKK_ARGTYPES
    t = name[0]
    herk, cplx = "herk" in name, t in "cz"
    a = np.zeros((4, 4), dtype=oa.NP_DTYPE[t], order="F")
    c = np.zeros((4, 4), dtype=oa.NP_DTYPE[t], order="F")
    bad_trans = "T" if herk else ("C" if cplx else "/")
    good_t = "C" if herk else "T"
    cases = [(1, "/", "N", 0, 0, 1, 1), (2, "U", bad_trans, 0, 0, 1, 1), (3, "U", "N", -1, 0, 1, 1), (4, "L", good_t, 0, -1, 1, 1),
             (7, "U", "N", 2, 0, 1, 2), (7, "L", good_t, 0, 2, 1, 1), (10, "U", "N", 2, 0, 2, 1), (10, "L", good_t, 2, 0, 1, 1)]
    label = (name[:-1].upper() + " ").encode()
    for (info, uplo, trans, n, k, lda, ldc) in cases:
        P.oracle_xerbla_expect(label, info)
        oa.call_rankk(getattr(L, name), name, uplo, trans, n, k, 1.0, a, lda, 1.0, c, ldc)
        assert P.oracle_xerbla_result() == 1, (name, info)
    # n == 0 returns before anything is touched
    assert oa.call_rankk(getattr(L, name), name, "U", "N", 0, 3, 1.0, a, 4, 0.0, c, 4) == 0


@pytest.mark.parametrize("name", oa.TRI_NAMES + oa.SYMM_NAMES + oa.R2K_NAMES)
def test_level3_error_exits_without_gpu(name):
    """?TRSM ?TRMM ?SYMM ?HEMM ?SYR2K ?HER2K blocks of xCHKE -- argument checking precedes any CUDA work; empty results
    return before anything is touched, with the reference's return values (blas/level3_impl.h:160,265,318,470)."""
    import level3_cases as lc
    L = eigen_b200.lib()
    lc.run_error_exits(oa.port(), getattr(L, name), name)
    t = name[0]
    z = np.zeros((4, 4), dtype=oa.NP_DTYPE[t], order="F")
    if name[1:] in ("trsm_", "trmm_"):
        assert oa.call_tri(getattr(L, name), name, "L", "U", "N", "N", 0, 3, 1.0, z, 4, z, 4) == (1 if "trmm" in name else 0)
    elif name[1:] in ("symm_", "hemm_"):
        assert oa.call_abc(getattr(L, name), name, "L", "U", 0, 3, 1.0, z, 4, z, 4, 0.0, z, 4) == 1
    else:
        assert oa.call_abc(getattr(L, name), name, "U", "N", 0, 3, 1.0, z, 4, z, 4, 0.0, z, 4) == 0
