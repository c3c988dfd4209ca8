"""Shared generators / checkers for ?potrf_ and ?getrf_ (TEST INFRASTRUCTURE; used by the CPU pin test and the GPU test).

The checks are LAPACK's own test ratios (TESTING/LIN dpot01 / dget01): the factors are multiplied back in long double
(oracle/hp_ref.c) and compared with the original matrix in units of eps * gauge."""
import numpy as np

import oracle_api as oa
from level3_cases import tol_for


def make_hpd(rng, t, n, ld=None):
    """Hermitian positive definite, entries O(1): M M^H / n + I, returned as an (ld x n) F-ordered array."""
    ld = n if ld is None else ld
    m = oa.rand_matrix(rng, t, n, n)[:n].astype(np.complex128 if t in "cz" else np.float64)
    h = m @ m.conj().T / max(n, 1) + np.eye(n)
    h = (h + h.conj().T) / 2
    out = oa.rand_matrix(rng, t, n, n, ld=ld)
    out[:n] = h.astype(oa.NP_DTYPE[t])
    return out


def poison_other_triangle(a, n, uplo):
    """NaN in the triangle ?potrf_ must not reference."""
    if n:
        a[:n][~oa.tri_mask(n, uplo)] = np.nan
    return a


def check_potrf(t, uplo, n, a0_full, a0, a, info):
    """a0_full: the Hermitian matrix; a0: what was passed in (maybe poisoned); a: the result."""
    assert info == 0, info
    assert a[n:].tobytes() == a0[n:].tobytes(), "ld padding was touched"
    if n == 0:
        return 0.0
    mask = oa.tri_mask(n, uplo)
    oc, o0 = a[:n][~mask], a0[:n][~mask]
    assert ((oc == o0) | (np.isnan(oc) & np.isnan(o0))).all(), "the unreferenced triangle was touched"
    f = np.asfortranarray(np.where(mask, a[:n], 0).astype(oa.NP_DTYPE[t]))
    assert np.all(np.diagonal(f).imag == 0) if t in "cz" else True
    assert np.all(np.diagonal(f).real > 0)
    zeros = np.zeros((n, n), dtype=oa.NP_DTYPE[t], order="F")
    if uplo in "Ll":
        ref, g = oa.hp_gemm(t, "N", "C", n, n, n, 1.0, f, n, f, n, 0.0, zeros, n)    # L L^H
    else:
        ref, g = oa.hp_gemm(t, "C", "N", n, n, n, 1.0, f, n, f, n, 0.0, zeros, n)    # U^H U
    want = np.asarray(a0_full[:n])
    with np.errstate(invalid="ignore", over="ignore"):
        ratio = (np.abs(ref - want) / (oa.EPS[t] * np.maximum(np.maximum(g, np.abs(want)), 1e-300)))[mask].max()
    assert ratio < tol_for(n), (t, uplo, n, ratio)
    return float(ratio)


def apply_ipiv(a, ipiv):
    """Row interchanges of xLASWP: for k ascending, swap rows k and ipiv[k]-1."""
    out = np.array(a, order="F")
    for k, p in enumerate(ipiv):
        p = int(p) - 1
        if p != k:
            out[[k, p]] = out[[p, k]]
    return out


def check_getrf(t, m, n, a0, a, ipiv, info, expect_info=0):
    assert info == expect_info, (info, expect_info)
    assert a[m:].tobytes() == a0[m:].tobytes(), "ld padding was touched"
    if m == 0 or n == 0:
        return 0.0
    k = min(m, n)
    assert len(ipiv) == k
    assert all(i + 1 <= int(p) <= m for i, p in enumerate(ipiv)), "pivot out of range"
    lu = a[:m]
    L = np.asfortranarray(np.tril(lu[:, :k], -1) + np.eye(m, k, dtype=lu.dtype))
    U = np.asfortranarray(np.triu(lu[:k, :]))
    if expect_info == 0:
        assert np.abs(np.tril(lu[:, :k], -1)).max(initial=0.0) <= 1.0 + 8 * oa.EPS[t], "partial pivoting bounds |L| by 1"
    zeros = np.zeros((m, n), dtype=oa.NP_DTYPE[t], order="F")
    ref, g = oa.hp_gemm(t, "N", "N", m, n, k, 1.0, L, m, U, k, 0.0, zeros, m)
    want = apply_ipiv(a0[:m], ipiv)
    with np.errstate(invalid="ignore", over="ignore"):
        ratio = (np.abs(ref - want) / (oa.EPS[t] * np.maximum(np.maximum(g, np.abs(want)), 1e-300))).max()
    assert ratio < tol_for(k), (t, m, n, ratio)
    return float(ratio)


POTRF_SIZES = (0, 1, 2, 3, 5, 9, 31, 32, 33, 64, 100, 257)
GETRF_SHAPES = ((0, 0), (1, 1), (2, 2), (3, 3), (5, 5), (9, 9), (16, 16), (17, 17), (33, 33), (64, 64), (100, 100), (257, 257),
                (9, 5), (100, 37), (300, 64), (257, 130))


def run_potrf_error_exits(P, fn, t):
    a = np.zeros((4, 4), dtype=oa.NP_DTYPE[t], order="F")
    label = (t.upper() + "POTRF").encode()
    for (info, uplo, n, lda) in [(1, "/", 0, 1), (2, "U", -1, 1), (4, "U", 2, 1)]:
        P.oracle_xerbla_expect(label, info)
        assert oa.call_potrf(fn, uplo, n, a, lda) == -info
        assert P.oracle_xerbla_result() == 1, (t, info)


def run_getrf_error_exits(P, fn, t):
    a = np.zeros((4, 4), dtype=oa.NP_DTYPE[t], order="F")
    label = (t.upper() + "GETRF").encode()
    for (info, m, n, lda) in [(1, -1, 0, 1), (2, 0, -1, 1), (4, 2, 1, 1)]:
        P.oracle_xerbla_expect(label, info)
        _, got = oa.call_getrf(fn, m, n, a, lda)
        assert got == -info
        assert P.oracle_xerbla_result() == 1, (t, info)
