"""Shared case generators / checkers for ?trsm_ ?trmm_ ?symm_ ?hemm_ ?syr2k_ ?her2k_ (TEST INFRASTRUCTURE).

Used by tests/test_oracle_pin_level3.py (oracle port vs the compiled reference, CPU) and tests/test_gpu_level3.py (the
sm_100a library, GPU).  Every check is the xBLAT3 one (blas/testing/dblat3.f DCHK2/DCHK3/DCHK5 + DMMCH): error measured
in units of eps * gauge against a long-double evaluation (oracle/hp_ref.c); for ?trsm_ the tester's residual form
op(A) * X  vs  alpha * B is used (dblat3.f:1100-1131)."""
import itertools

import numpy as np

import oracle_api as oa


def tol_for(k):
    """netlib's threshold is 16 gauge units for k <= 9; chained sums drift like sqrt(k) * eps (see test_gpu_rankk.py)."""
    return max(16.0, 4.0 * np.sqrt(max(k, 1)))


def scalars(name):
    """(alphas, betas) of the netlib .dat grids; real for the real types and for the REAL beta of ?her2k_."""
    t = name[0]
    if t in "sd":
        return [0.0, 1.0, 0.7], [0.0, 1.0, 1.3]
    if "her2k" in name:
        return [0.0, 1.0, 0.7 - 0.9j], [0.0, 1.0, 1.3]
    return [0.0, 1.0, 0.7 - 0.9j], [0.0, 1.0, 1.3 - 1.1j]


def legal_trans(name):
    t = name[0]
    if "her2k" in name:
        return "NC"
    if "syr2k" in name:
        return "NT" if t in "cz" else "NTC"
    return "NTC"


def _ratio(t, got, ref, g, mask=None):
    with np.errstate(over="ignore", invalid="ignore"):
        err = np.abs(got - ref) / (oa.EPS[t] * np.maximum(g, 1e-300))
    if mask is not None:
        err = err[mask]
    return float(err.max()) if err.size else 0.0


# ---- ?trsm_ / ?trmm_ ----------------------------------------------------------------------------------------------
def tri_inputs(rng, name, side, m, n, lda_pad=1, ldb_pad=1, well_conditioned=True):
    t = name[0]
    na = m if side in "Ll" else n
    if well_conditioned:
        a = oa.make_triangular(rng, t, na, na + lda_pad)
    else:
        a = oa.rand_matrix(rng, t, na, na, ld=na + lda_pad)
    b = oa.rand_matrix(rng, t, m, n, ld=m + ldb_pad)
    return a, b


def check_tri(name, side, uplo, trans, diag, m, n, alpha, a, b0, b, tol=None):
    """b0 = B on entry, b = B on return.  Padding untouched; result within tol gauge units."""
    t = name[0]
    na = m if side in "Ll" else n
    tol = tol_for(na) if tol is None else tol
    assert b[m:].tobytes() == b0[m:].tobytes(), "ld padding of B was touched"
    if m == 0 or n == 0:
        return 0.0
    T = oa.dense_triangular(a, na, uplo, diag)
    X = np.asfortranarray(b[:m].copy())
    B0 = np.asfortranarray(b0[:m].copy())
    zeros = np.zeros((m, n), dtype=oa.NP_DTYPE[t], order="F")
    left = side in "Ll"
    if "trmm" in name:
        if left:
            ref, g = oa.hp_gemm(t, trans, "N", m, n, m, alpha, T, na, B0, m, 0.0, zeros, m)
        else:
            ref, g = oa.hp_gemm(t, "N", trans, m, n, n, alpha, B0, m, T, na, 0.0, zeros, m)
        r = _ratio(t, X, ref, g)
    else:
        # residual form: op(T) * X (or X * op(T)) against alpha * B0
        if left:
            ref, g = oa.hp_gemm(t, trans, "N", m, n, m, 1.0, T, na, X, m, 0.0, zeros, m)
        else:
            ref, g = oa.hp_gemm(t, "N", trans, m, n, n, 1.0, X, m, T, na, 0.0, zeros, m)
        want = np.asarray(alpha, dtype=np.complex128 if t in "cz" else np.float64) * B0.astype(ref.dtype)
        r = _ratio(t, ref, want, np.maximum(g, np.abs(want)))
    assert r < tol, (name, side, uplo, trans, diag, m, n, alpha, r)
    return r


# ---- ?symm_ / ?hemm_ ----------------------------------------------------------------------------------------------
def check_symm(name, side, uplo, m, n, alpha, beta, a, b, c0, c, tol=None):
    t = name[0]
    herm = "hemm" in name
    na = m if side in "Ll" else n
    tol = tol_for(na) if tol is None else tol
    assert c[m:].tobytes() == c0[m:].tobytes(), "ld padding of C was touched"
    if m == 0 or n == 0:
        return 0.0
    S = oa.dense_symmetric(a, na, uplo, herm)
    Cin = np.asfortranarray(c0[:m]) if beta != 0 else np.zeros((m, n), dtype=c0.dtype, order="F")
    Bm = np.asfortranarray(b[:m])
    if side in "Ll":
        ref, g = oa.hp_gemm(t, "N", "N", m, n, m, alpha, S, na, Bm, m, beta, Cin, m)
    else:
        ref, g = oa.hp_gemm(t, "N", "N", m, n, n, alpha, Bm, m, S, na, beta, Cin, m)
    r = _ratio(t, c[:m], ref, g)
    assert r < tol, (name, side, uplo, m, n, alpha, beta, r)
    return r


# ---- ?syr2k_ / ?her2k_ --------------------------------------------------------------------------------------------
def check_r2k(name, uplo, trans, n, k, alpha, beta, a, b, c0, c, tol=None, exact_real_diagonal=False):
    t = name[0]
    her = "her2k" in name
    tol = tol_for(2 * k) if tol is None else tol
    mask = oa.tri_mask(n, uplo)
    assert c[n:].tobytes() == c0[n:].tobytes(), "ld padding of C was touched"
    oc, o0 = c[:n][~mask], c0[:n][~mask]
    assert ((oc == o0) | (np.isnan(oc) & np.isnan(o0))).all(), "the unreferenced triangle was touched"
    if n == 0:
        return 0.0
    other = "C" if her else "T"
    ta, tb = ("N", other) if trans in "Nn" else (other, "N")
    lda, ldb = a.shape[0], b.shape[0]
    Cin = np.asfortranarray(c0[:n]) if beta != 0 else np.zeros((n, n), dtype=c0.dtype, order="F")
    zeros = np.zeros((n, n), dtype=c0.dtype, order="F")
    r1, g1 = oa.hp_gemm(t, ta, tb, n, n, k, alpha, a, lda, b, ldb, beta, Cin, n)
    alpha2 = np.conj(alpha) if her else alpha
    r2, g2 = oa.hp_gemm(t, ta, tb, n, n, k, alpha2, b, ldb, a, lda, 0.0, zeros, n)
    ref, g = r1 + r2, g1 + g2
    got = c[:n].astype(ref.dtype)
    untouched = (beta == 1.0) and (k == 0 or alpha == 0.0)
    if her and not untouched:
        ref[np.arange(n), np.arange(n)] = ref[np.arange(n), np.arange(n)].real
        if exact_real_diagonal:
            assert np.all(np.diagonal(c[:n]).imag == 0), "Hermitian diagonal must be exactly real"
    r = _ratio(t, got, ref, g, mask)
    assert r < tol, (name, uplo, trans, n, k, alpha, beta, r)
    return r


SMALL_DIMS = (0, 1, 2, 3, 5, 9)


def sweep_tri(fn, name, rng, dims=SMALL_DIMS, extra=((35, 7), (7, 35), (40, 70))):
    """DCHK3 (dblat3.f:925-1252): every side/uplo/trans/diag, dims from the .dat grid plus sizes beyond one leaf."""
    alphas, _ = scalars(name)
    worst = 0.0
    shapes = list(itertools.product(dims, repeat=2)) + list(extra)
    for (m, n) in shapes:
        for side, uplo, trans, diag in itertools.product("LR", "UL", "NTC", "UN"):
            for alpha in alphas:
                a, b0 = tri_inputs(rng, name, side, m, n)
                a0 = a.copy(order="F")
                b = b0.copy(order="F")
                ret = oa.call_tri(fn, name, side, uplo, trans, diag, m, n, alpha, a, a.shape[0], b, b.shape[0])
                assert ret == (1 if "trmm" in name else 0), (name, ret)   # level3_impl.h:265,283 vs :160,177
                assert a.tobytes() == a0.tobytes(), "A was modified"
                worst = max(worst, check_tri(name, side, uplo, trans, diag, m, n, alpha, a, b0, b))
    return worst


def sweep_symm(fn, name, rng, dims=SMALL_DIMS, extra=((35, 7), (7, 35), (40, 70))):
    """DCHK2 (dblat3.f:677-923)."""
    t = name[0]
    alphas, betas = scalars(name)
    worst = 0.0
    for (m, n) in list(itertools.product(dims, repeat=2)) + list(extra):
        for side, uplo in itertools.product("LR", "UL"):
            na = m if side == "L" else n
            a = oa.rand_matrix(rng, t, na, na, ld=na + 1)
            if na:   # poison the unreferenced triangle: it must never be read
                poison = ~oa.tri_mask(na, uplo)
                a[:na][poison] = np.nan
            b = oa.rand_matrix(rng, t, m, n, ld=m + 1)
            for alpha, beta in itertools.product(alphas, betas):
                c0 = oa.rand_matrix(rng, t, m, n, ld=m + 1)
                if beta == 0 and m:
                    c0[:m] = np.nan   # beta == 0 must not read C
                c = c0.copy(order="F")
                oa.call_abc(fn, name, side, uplo, m, n, alpha, a, na + 1, b, m + 1, beta, c, m + 1)
                worst = max(worst, check_symm(name, side, uplo, m, n, alpha, beta, a, b, c0, c))
    return worst


def sweep_r2k(fn, name, rng, dims=SMALL_DIMS, extra=((35, 7), (7, 35), (70, 40)), exact_real_diagonal=False):
    """DCHK5 (dblat3.f:1549-1887)."""
    t = name[0]
    alphas, betas = scalars(name)
    worst = 0.0
    for (n, k) in list(itertools.product(dims, repeat=2)) + list(extra):
        for uplo, trans in itertools.product("UL", legal_trans(name)):
            ra, ca = (n, k) if trans == "N" else (k, n)
            a = oa.rand_matrix(rng, t, ra, ca, ld=ra + 1)
            b = oa.rand_matrix(rng, t, ra, ca, ld=ra + 1)
            for alpha, beta in itertools.product(alphas, betas):
                c0 = oa.rand_matrix(rng, t, n, n, ld=n + 1)
                if beta == 0 and n:
                    c0[:n][oa.tri_mask(n, uplo)] = np.nan
                c = c0.copy(order="F")
                oa.call_abc(fn, name, uplo, trans, n, k, alpha, a, ra + 1, b, ra + 1, beta, c, n + 1)
                worst = max(worst, check_r2k(name, uplo, trans, n, k, alpha, beta, a, b, c0, c,
                                             exact_real_diagonal=exact_real_diagonal))
    return worst


def tri_error_cases():
    """The ?TRSM / ?TRMM block of xCHKE: (info, side, uplo, trans, diag, m, n, lda, ldb)."""
    return [(1, "/", "U", "N", "N", 0, 0, 1, 1), (2, "L", "/", "N", "N", 0, 0, 1, 1), (3, "L", "U", "/", "N", 0, 0, 1, 1),
            (4, "L", "U", "N", "/", 0, 0, 1, 1), (5, "L", "U", "N", "N", -1, 0, 1, 1), (5, "R", "L", "T", "U", -1, 0, 1, 1),
            (6, "L", "U", "N", "N", 0, -1, 1, 1), (6, "R", "L", "T", "N", 0, -1, 1, 1), (9, "L", "U", "N", "N", 2, 0, 1, 2),
            (9, "R", "U", "N", "N", 0, 2, 1, 1), (9, "L", "L", "T", "N", 2, 0, 1, 2), (11, "L", "U", "N", "N", 2, 0, 2, 1),
            (11, "R", "L", "N", "N", 2, 0, 1, 1)]


def symm_error_cases():
    """(info, side, uplo, m, n, lda, ldb, ldc)."""
    return [(1, "/", "U", 0, 0, 1, 1, 1), (2, "L", "/", 0, 0, 1, 1, 1), (3, "L", "U", -1, 0, 1, 1, 1), (3, "R", "L", -1, 0, 1, 1, 1),
            (4, "L", "U", 0, -1, 1, 1, 1), (4, "R", "L", 0, -1, 1, 1, 1), (7, "L", "U", 2, 0, 1, 2, 2), (7, "R", "U", 0, 2, 1, 1, 1),
            (9, "L", "U", 2, 0, 2, 1, 2), (9, "R", "L", 2, 0, 1, 1, 2), (12, "L", "U", 2, 0, 2, 2, 1), (12, "R", "L", 2, 0, 1, 2, 1)]


def r2k_error_cases(name):
    """(info, uplo, trans, n, k, lda, ldb, ldc)."""
    t = name[0]
    her = "her2k" in name
    bad = "T" if her else ("C" if t in "cz" else "/")
    good = "C" if her else "T"
    return [(1, "/", "N", 0, 0, 1, 1, 1), (2, "U", bad, 0, 0, 1, 1, 1), (3, "U", "N", -1, 0, 1, 1, 1), (3, "L", good, -1, 0, 1, 1, 1),
            (4, "U", "N", 0, -1, 1, 1, 1), (4, "L", good, 0, -1, 1, 1, 1), (7, "U", "N", 2, 0, 1, 2, 2), (7, "U", good, 0, 2, 1, 2, 1),
            (9, "U", "N", 2, 0, 2, 1, 2), (9, "L", good, 0, 2, 2, 1, 1), (12, "U", "N", 2, 0, 2, 2, 1), (12, "L", good, 2, 0, 1, 1, 1)]


def run_error_exits(P, fn, name):
    """Every illegal call must reach xerbla_ with the expected routine name and info, and touch nothing."""
    t = name[0]
    a = np.zeros((4, 4), dtype=oa.NP_DTYPE[t], order="F")
    b = np.zeros((4, 4), dtype=oa.NP_DTYPE[t], order="F")
    c = np.zeros((4, 4), dtype=oa.NP_DTYPE[t], order="F")
    label = name[:-1].upper().ljust(6).encode()
    if name[1:] in ("trsm_", "trmm_"):
        for (info, side, uplo, trans, diag, m, n, lda, ldb) in tri_error_cases():
            P.oracle_xerbla_expect(label, info)
            oa.call_tri(fn, name, side, uplo, trans, diag, m, n, 1.0, a, lda, b, ldb)
            assert P.oracle_xerbla_result() == 1, (name, info)
    elif name[1:] in ("symm_", "hemm_"):
        for (info, side, uplo, m, n, lda, ldb, ldc) in symm_error_cases():
            P.oracle_xerbla_expect(label, info)
            oa.call_abc(fn, name, side, uplo, m, n, 1.0, a, lda, b, ldb, 1.0, c, ldc)
            assert P.oracle_xerbla_result() == 1, (name, info)
    else:
        for (info, uplo, trans, n, k, lda, ldb, ldc) in r2k_error_cases(name):
            P.oracle_xerbla_expect(label, info)
            oa.call_abc(fn, name, uplo, trans, n, k, 1.0, a, lda, b, ldb, 1.0, c, ldc)
            assert P.oracle_xerbla_result() == 1, (name, info)
