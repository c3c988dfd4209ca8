"""?trsm_ ?trmm_ ?symm_ ?hemm_ ?syr2k_ ?her2k_ on the sm_100a library -- run with -m gpu on a B200.

Reference semantics: blas/level3_impl.h:78-355, :437-562, :631-700.  The checks are the xBLAT3 ones (level3_cases.py)
against the long-double oracle; the same sweeps pass on the compiled reference in tests/test_oracle_pin_level3.py."""
import ctypes as C
import itertools

import numpy as np
import pytest

import eigen_b200
import level3_cases as lc
import oracle_api as oa

pytestmark = pytest.mark.gpu
P = oa.port()


@pytest.fixture(scope="module")
def L():
    lib = eigen_b200.require_device()
    yield lib
    lib.b200blas_set_variant(0)


@pytest.mark.parametrize("name", oa.TRI_NAMES)
def test_tri_xblat3_sweep(L, name):
    lc.sweep_tri(getattr(L, name), name, np.random.default_rng(31))


@pytest.mark.parametrize("name", oa.SYMM_NAMES)
def test_symm_xblat3_sweep(L, name):
    lc.sweep_symm(getattr(L, name), name, np.random.default_rng(32))


@pytest.mark.parametrize("name", oa.R2K_NAMES)
def test_r2k_xblat3_sweep(L, name):
    lc.sweep_r2k(getattr(L, name), name, np.random.default_rng(33), exact_real_diagonal=True)


@pytest.mark.parametrize("name", oa.TRI_NAMES + oa.SYMM_NAMES + oa.R2K_NAMES)
def test_error_exits(L, name):
    lc.run_error_exits(P, getattr(L, name), name)


@pytest.mark.parametrize("name", oa.TRI_NAMES)
def test_tri_larger_shapes(L, name):
    """Several recursion levels, ragged leaves, off-diagonal blocks large enough for the tensor-pipe kernels, both the
    LU shape (small unit-lower triangle, long B; PartialPivLU.h:490) and the LLT shape (solve on the right; LLT.h:356)."""
    rng = np.random.default_rng(41)
    cases = [("L", "L", "N", "U", 256, 1500), ("R", "L", "C", "N", 1500, 128), ("L", "U", "N", "N", 777, 300),
             ("R", "U", "T", "N", 300, 777), ("L", "L", "T", "N", 1030, 257), ("R", "L", "N", "U", 129, 1030)]
    alphas, _ = lc.scalars(name)
    for i, (side, uplo, trans, diag, m, n) in enumerate(cases):
        a, b0 = lc.tri_inputs(rng, name, side, m, n, lda_pad=i % 3, ldb_pad=(i + 1) % 3)
        b = b0.copy(order="F")
        oa.call_tri(getattr(L, name), name, side, uplo, trans, diag, m, n, alphas[2], a, a.shape[0], b, b.shape[0])
        lc.check_tri(name, side, uplo, trans, diag, m, n, alphas[2], a, b0, b)


@pytest.mark.parametrize("name", oa.SYMM_NAMES + oa.R2K_NAMES)
def test_abc_larger_shapes(L, name):
    t = name[0]
    rng = np.random.default_rng(43)
    alphas, betas = lc.scalars(name)
    if name in oa.SYMM_NAMES:
        for i, (side, uplo, m, n) in enumerate([("L", "U", 777, 300), ("R", "L", 300, 777), ("L", "L", 1030, 129), ("R", "U", 129, 1030)]):
            na = m if side == "L" else n
            a = oa.rand_matrix(rng, t, na, na, ld=na + i % 2)
            a[:na][~oa.tri_mask(na, uplo)] = np.nan
            b = oa.rand_matrix(rng, t, m, n, ld=m + 1)
            c0 = oa.rand_matrix(rng, t, m, n, ld=m + 2)
            c = c0.copy(order="F")
            oa.call_abc(getattr(L, name), name, side, uplo, m, n, alphas[2], a, a.shape[0], b, m + 1, betas[1 + i % 2], c, m + 2)
            lc.check_symm(name, side, uplo, m, n, alphas[2], betas[1 + i % 2], a, b, c0, c)
    else:
        for i, (n, k) in enumerate([(300, 513), (777, 40), (1030, 700)]):
            for uplo, trans in itertools.product("UL", lc.legal_trans(name)[:2]):
                ra, ca = (n, k) if trans == "N" else (k, n)
                a = oa.rand_matrix(rng, t, ra, ca, ld=ra + i % 2)
                b = oa.rand_matrix(rng, t, ra, ca, ld=ra + 1)
                c0 = oa.rand_matrix(rng, t, n, n, ld=n + 2)
                c = c0.copy(order="F")
                oa.call_abc(getattr(L, name), name, uplo, trans, n, k, alphas[2], a, a.shape[0], b, ra + 1, betas[1 + i % 2], c, n + 2)
                lc.check_r2k(name, uplo, trans, n, k, alphas[2], betas[1 + i % 2], a, b, c0, c, exact_real_diagonal=True)


def test_dtrsm_device_resident_at_scale(L):
    """Device pointers through the F77 entry, m = n = 8192: residual on sampled rows against the long-double oracle, and
    the solve undone by dtrmm returns alpha^2 * B (size-independent round-trip property)."""
    import torch
    m = n = 8192
    rng = np.random.default_rng(47)
    a = oa.make_triangular(rng, "d", m, m)
    b0 = oa.rand_matrix(rng, "d", m, n)
    dA = torch.from_numpy(np.ascontiguousarray(a.T)).cuda()      # column-major m x m as a (cols, rows) tensor
    dB = torch.from_numpy(np.ascontiguousarray(b0.T)).cuda()
    ints = [C.c_int(v) for v in (m, n, m, m)]
    al = C.c_double(0.5)
    r = L.dtrsm_(b"L", b"L", b"N", b"N", C.byref(ints[0]), C.byref(ints[1]), C.byref(al), C.c_void_p(dA.data_ptr()), C.byref(ints[2]),
                 C.c_void_p(dB.data_ptr()), C.byref(ints[3]))
    assert r == 0
    x = np.asfortranarray(dB.cpu().numpy().T)
    rows = np.sort(rng.choice(m, size=24, replace=False)).astype(np.int32)
    T = oa.dense_triangular(a, m, "L", "N")
    zeros = np.zeros((m, n), order="F")
    ref, g = oa.hp_gemm("d", "N", "N", m, n, m, 1.0, T, m, x, m, 0.0, zeros, m, rows=rows)
    want = 0.5 * b0[rows]
    ratio = (np.abs(ref - want) / (oa.EPS["d"] * np.maximum(g, np.abs(want)))).max()
    assert ratio < lc.tol_for(m), ratio
    r = L.dtrmm_(b"L", b"L", b"N", b"N", C.byref(ints[0]), C.byref(ints[1]), C.byref(al), C.c_void_p(dA.data_ptr()), C.byref(ints[2]),
                 C.c_void_p(dB.data_ptr()), C.byref(ints[3]))
    assert r == 1
    back = dB.cpu().numpy().T
    err = np.abs(back - 0.25 * b0).max()
    assert err < 1e-11, err
