"""?syrk_ / ?herk_ on the sm_100a library against the CPU oracle -- run with -m gpu on a B200.

Reference semantics: blas/level3_impl.h:357-433 (syrk), :564-627 (herk).  The kernels are the GEMM kernels with B := A
and a triangular tile mask, so the bar is the same: netlib gauge ratio < 16 against the long-double reference, only the
referenced triangle written, ld padding and the other triangle bit-identical, Hermitian diagonal exactly real."""
import ctypes as C
import itertools

import numpy as np
import pytest

import eigen_b200
import oracle_api as oa

pytestmark = pytest.mark.gpu
P = oa.port()
NAMES = oa.RANKK_NAMES


@pytest.fixture(scope="module")
def L():
    lib = eigen_b200.require_device()
    yield lib
    lib.b200blas_set_variant(0)


def _params(name):
    t = name[0]
    cplx, herk = t in "cz", "herk" in name
    transes = "NC" if herk else ("NT" if cplx else "NTC")
    if herk or not cplx:
        alphas, betas = [0.0, 1.0, 0.7], [0.0, 1.0, 1.3]
    else:
        alphas, betas = [0.0, 1.0, 0.7 - 0.9j], [0.0, 1.0, 1.3 - 1.1j]
    return transes, alphas, betas


def _check(name, uplo, trans, n, k, alpha, beta, A, lda, C0, c, ldc, tol=None):
    t = name[0]
    # netlib's threshold (16 gauge units) is meant for k <= 9.  On the diagonal of a rank-k update every term is
    # positive, so the gauge equals the value and any chained summation drifts like sqrt(k) * eps (deterministic bound
    # k * eps); a wrong element is off by ~1 / eps gauge units, so 4 * sqrt(k) still separates the two cleanly.
    tol = max(16.0, 4.0 * np.sqrt(k)) if tol is None else tol
    m = oa.tri_mask(n, uplo)
    assert c[n:].tobytes() == C0[n:].tobytes(), "ld padding of C was touched"
    other_c, other_0 = c[:n][~m], C0[:n][~m]
    same = (other_c == other_0) | (np.isnan(other_c) & np.isnan(other_0))
    assert same.all(), "the unreferenced triangle was touched"
    Cin = C0 if beta != 0 else np.zeros_like(C0)
    ref, g = oa.hp_rankk(name, uplo, trans, n, k, alpha, A, lda, beta, Cin, ldc)
    if "herk" in name and not (beta == 1.0 and (k == 0 or alpha == 0.0)):
        assert np.all(np.diagonal(c[:n]).imag == 0), "Hermitian diagonal must be exactly real"
        ref[np.arange(n), np.arange(n)] = ref[np.arange(n), np.arange(n)].real
    ratio = (np.abs(c[:n] - ref)[m] / (oa.EPS[t] * np.maximum(g, 1e-300)[m])).max() if n else 0.0
    assert ratio < tol, (name, uplo, trans, n, k, alpha, beta, ratio, eigen_b200.last_variant())


@pytest.mark.parametrize("name", NAMES)
def test_xblat3_style_sweep(L, name):
    """DCHK4 of the netlib tester (blas/testing/dblat3.f:1254-1547): dims 0 1 2 3 5 9, ld = dim+1, both triangles, every
    legal trans, alpha/beta from the .dat files; run under every kernel variant."""
    t = name[0]
    rng = np.random.default_rng(23)
    transes, alphas, betas = _params(name)
    for v in (["simt", "dmma"] if t in "dz" else ["simt", "tf32x3"]):
        L.b200blas_set_variant(eigen_b200.VARIANT[v])
        for n, k in itertools.product((0, 1, 2, 3, 5, 9), repeat=2):
            for uplo, trans in itertools.product("UL", transes):
                ra, ca = (n, k) if trans == "N" else (k, n)
                A = oa.rand_matrix(rng, t, ra, ca, ld=ra + 1)
                A0 = A.copy(order="F")
                for alpha, beta in itertools.product(alphas, betas):
                    C0 = oa.rand_matrix(rng, t, n, n, ld=n + 1)
                    if beta == 0:
                        C0[:n][oa.tri_mask(n, uplo)] = np.nan   # beta == 0 must not read C
                    c = C0.copy(order="F")
                    assert oa.call_rankk(getattr(L, name), name, uplo, trans, n, k, alpha, A, ra + 1, beta, c, n + 1) == 0
                    assert A.tobytes() == A0.tobytes()
                    if n:
                        _check(name, uplo, trans, n, k, alpha, beta, A, ra + 1, C0, c, n + 1)
    L.b200blas_set_variant(0)


@pytest.mark.parametrize("name", NAMES)
def test_error_exits(L, name):
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


@pytest.mark.parametrize("name", NAMES)
def test_random_shapes(L, name):
    """Ragged sizes that cross tile boundaries, diagonal tiles, k-slices and the pageable staging ring."""
    t = name[0]
    rng = np.random.default_rng(29)
    transes, alphas, betas = _params(name)
    shapes = [(129, 64), (300, 513), (777, 40), (1500, 300), (1030, 2100)]
    for i, (n, k) in enumerate(shapes):
        for uplo, trans in itertools.product("UL", transes):
            ra, ca = (n, k) if trans == "N" else (k, n)
            A = oa.rand_matrix(rng, t, ra, ca, ld=ra + (i % 2))
            C0 = oa.rand_matrix(rng, t, n, n, ld=n + 2)
            alpha, beta = alphas[2], betas[1 + (i % 2)]
            c = C0.copy(order="F")
            assert oa.call_rankk(getattr(L, name), name, uplo, trans, n, k, alpha, A, A.shape[0], beta, c, n + 2) == 0
            _check(name, uplo, trans, n, k, alpha, beta, A, A.shape[0], C0, c, n + 2)


def test_dsyrk_equals_dgemm_on_the_triangle_at_scale(L):
    """Size-independent property at full scale, device-resident: the rank-k update is the GEMM kernel with B := A^T and a
    mask, so dsyrk(n = 8192, k = 4096) must equal dgemm(A, A^T) BIT FOR BIT on the triangle and leave the rest alone."""
    import torch
    n, k = 8192, 4096
    g = torch.Generator(device="cuda").manual_seed(3)
    A = torch.rand(k, n, dtype=torch.float64, device="cuda", generator=g) * 2 - 1      # column-major n x k
    C0 = torch.rand(n, n, dtype=torch.float64, device="cuda", generator=g)
    Cg = C0.clone()
    assert eigen_b200.gemm_dev("d", "N", "T", n, n, k, -1.0, A, n, A, n, 1.0, Cg, n) == 0
    for uplo in "LU":
        Cs = C0.clone()
        ints = [C.c_int(v) for v in (n, k, n, n)]
        al, be = C.c_double(-1.0), C.c_double(1.0)
        r = L.dsyrk_(uplo.encode(), b"N", C.byref(ints[0]), C.byref(ints[1]), C.byref(al), C.c_void_p(A.data_ptr()), C.byref(ints[2]),
                     C.byref(be), C.c_void_p(Cs.data_ptr()), C.byref(ints[3]))
        assert r == 0
        torch.cuda.synchronize()
        # tensor (col, row): element (i, j) at [j, i]; lower triangle i >= j  <=>  tensor upper triangle incl. diagonal
        tri = torch.triu(torch.ones(n, n, dtype=torch.bool, device="cuda")) if uplo == "L" else torch.tril(torch.ones(n, n, dtype=torch.bool, device="cuda"))
        assert torch.equal(Cs[tri], Cg[tri])
        assert torch.equal(Cs[~tri], C0[~tri])
