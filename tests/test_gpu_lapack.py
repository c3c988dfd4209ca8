"""?potrf_ / ?getrf_ on the sm_100a library -- run with -m gpu on a B200.

Reference semantics: lapack/cholesky.cpp:14-38 (LLT.h:299-360) and lapack/lu.cpp:14-42 (PartialPivLU.h:361-496).  Checks are
LAPACK's factor-product ratios against the long-double oracle (lapack_cases.py); the same checks pass on the compiled
reference in tests/test_oracle_pin_lapack.py.  Pivot sequences are compared with the oracle port's."""
import ctypes as C

import numpy as np
import pytest

import eigen_b200
import lapack_cases as lp
import oracle_api as oa

pytestmark = pytest.mark.gpu
P = oa.port()


@pytest.fixture(scope="module")
def L():
    return eigen_b200.require_device()


@pytest.mark.parametrize("t", list("sdcz"))
def test_potrf_sweep(L, t):
    rng = np.random.default_rng(3)
    for n in lp.POTRF_SIZES + (1030,):
        for uplo in "LU":
            full = lp.make_hpd(rng, t, n, ld=n + 1)
            a0 = lp.poison_other_triangle(full.copy(order="F"), n, uplo)
            a = a0.copy(order="F")
            info = oa.call_potrf(getattr(L, t + "potrf_"), uplo, n, a, n + 1)
            lp.check_potrf(t, uplo, n, full, a0, a, info)


@pytest.mark.parametrize("t", list("sdcz"))
def test_getrf_sweep_and_pivots_match_the_oracle(L, t):
    rng = np.random.default_rng(4)
    for (m, n) in lp.GETRF_SHAPES + ((1030, 1030), (1500, 700), (37, 100), (130, 257)):
        a0 = oa.rand_matrix(rng, t, m, n, ld=m + 1)
        a = a0.copy(order="F")
        ipiv, info = oa.call_getrf(getattr(L, t + "getrf_"), m, n, a, m + 1)
        lp.check_getrf(t, m, n, a0, a, ipiv, info)
        if m >= n and max(m, n) <= 300:   # the reference leaves the columns right of a wide matrix's square part unfinished
            b = a0.copy(order="F")
            opiv, oinfo = oa.call_getrf(getattr(P, "oracle_%sgetrf_" % t), m, n, b, m + 1)
            assert oinfo == info
            assert np.array_equal(opiv, ipiv), (t, m, n)
            assert np.abs(a[:m] - b[:m]).max(initial=0.0) <= 4096 * oa.EPS[t] * max(1.0, float(np.abs(b[:m]).max(initial=0.0)))


@pytest.mark.parametrize("t,m,n", [("d", 9000, 70), ("s", 9000, 70), ("z", 9000, 40), ("d", 16384, 33), ("c", 20000, 48), ("d", 140000, 40),
                                   ("z", 70000, 20)])
def test_getrf_tall_panels(L, t, m, n):
    """Panel heights on both sides of every leaf-kernel boundary (lapack.cu RegPanel: <= 8192 rows wide register leaf, <= 16384
    narrow one with two rows per thread, above that the cooperative slab kernel, and past ~120k rows of doubles its
    global-memory slab -- ADVICE r1: a 200000 x 64 LU must not fail); pivots identical to the oracle port's."""
    rng = np.random.default_rng(14)
    a0 = oa.rand_matrix(rng, t, m, n, ld=m + 1)
    a = a0.copy(order="F")
    ipiv, info = oa.call_getrf(getattr(L, t + "getrf_"), m, n, a, m + 1)
    lp.check_getrf(t, m, n, a0, a, ipiv, info)
    b = a0.copy(order="F")
    opiv, oinfo = oa.call_getrf(getattr(P, "oracle_%sgetrf_" % t), m, n, b, m + 1)
    assert oinfo == info
    assert np.array_equal(opiv, ipiv), (t, m, n)


@pytest.mark.parametrize("t", list("sdcz"))
def test_failure_reports(L, t):
    rng = np.random.default_rng(6)
    for n, k in ((9, 4), (100, 57), (257, 200), (1030, 700)):
        a = lp.make_hpd(rng, t, n)
        a[k, k] = -1.0
        assert oa.call_potrf(getattr(L, t + "potrf_"), "L", n, a.copy(order="F"), n) == k + 1
        assert oa.call_potrf(getattr(L, t + "potrf_"), "U", n, a.copy(order="F"), n) == k + 1
    for (m, n, k) in ((9, 9, 4), (100, 100, 57), (257, 130, 100), (1030, 1030, 700)):
        a0 = oa.rand_matrix(rng, t, m, n)
        a0[:, k] = 0
        a = a0.copy(order="F")
        ipiv, info = oa.call_getrf(getattr(L, t + "getrf_"), m, n, a, m)
        lp.check_getrf(t, m, n, a0, a, ipiv, info, expect_info=k + 1)


@pytest.mark.parametrize("t", list("sdcz"))
def test_error_exits(L, t):
    lp.run_potrf_error_exits(P, getattr(L, t + "potrf_"), t)
    lp.run_getrf_error_exits(P, getattr(L, t + "getrf_"), t)


def test_device_resident_factorizations_at_scale(L):
    """n = 8192, matrix already in HBM (device pointer through the F77 entry): sampled rows of L L^T and of P A = L U
    against the long-double oracle."""
    import torch
    n = 8192
    rng = np.random.default_rng(8)
    rows = np.sort(rng.choice(n, size=16, replace=False)).astype(np.int32)
    zeros = np.zeros((n, n), order="F")
    # Cholesky
    full = lp.make_hpd(rng, "d", n)
    dA = torch.from_numpy(np.ascontiguousarray(full.T)).cuda()
    info = C.c_int(-7)
    nn = C.c_int(n)
    assert L.dpotrf_(b"L", C.byref(nn), C.c_void_p(dA.data_ptr()), C.byref(nn), C.byref(info)) == 0
    assert info.value == 0
    f = np.asfortranarray(np.tril(dA.cpu().numpy().T))
    ref, g = oa.hp_gemm("d", "N", "C", n, n, n, 1.0, f, n, f, n, 0.0, zeros, n, rows=rows)
    mask = np.arange(n)[None, :] <= rows[:, None]
    ratio = (np.abs(ref - full[rows]) / (oa.EPS["d"] * np.maximum(g, np.abs(full[rows]))))[mask].max()
    assert ratio < lp.tol_for(n), ratio
    # LU
    a0 = oa.rand_matrix(rng, "d", n, n)
    dA = torch.from_numpy(np.ascontiguousarray(a0.T)).cuda()
    ipiv = np.zeros(n, dtype=np.int32)
    assert L.dgetrf_(C.byref(nn), C.byref(nn), C.c_void_p(dA.data_ptr()), C.byref(nn), ipiv.ctypes.data_as(C.POINTER(C.c_int)), C.byref(info)) == 0
    assert info.value == 0
    lu = np.asfortranarray(dA.cpu().numpy().T)
    Lm = np.asfortranarray(np.tril(lu, -1) + np.eye(n))
    U = np.asfortranarray(np.triu(lu))
    assert np.abs(np.tril(lu, -1)).max() <= 1.0 + 1e-12
    ref, g = oa.hp_gemm("d", "N", "N", n, n, n, 1.0, Lm, n, U, n, 0.0, zeros, n, rows=rows)
    want = lp.apply_ipiv(a0, ipiv)[rows]
    ratio = (np.abs(ref - want) / (oa.EPS["d"] * np.maximum(g, np.abs(want)))).max()
    assert ratio < lp.tol_for(n), ratio
