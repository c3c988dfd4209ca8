"""Pins the rank-k part of the CPU oracle (oracle/rankk_port.c: ?syrk_ / ?herk_) against the reference's
own blas/ library (oracle/_ref/libeigen_blas_ref.so, blas/level3_impl.h:357-433, 564-627) -- CPU only."""
import itertools

import numpy as np
import pytest

import oracle_api as oa

P = oa.port()
needs_ref = pytest.mark.skipif(not oa.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
NAMES = oa.RANKK_NAMES


def _params(name):
    t = name[0]
    cplx, herk = t in "cz", "herk" in name
    transes = "NC" if herk else ("NT" if cplx else "NTC")
    if herk or not cplx:
        alphas, betas = [0.0, 1.0, 0.7], [0.0, 1.0, 1.3]
    else:
        alphas, betas = [0.0, 1.0, 0.7 - 0.9j], [0.0, 1.0, 1.3 - 1.1j]   # {c,z}blat3.dat:12-14
    return transes, alphas, betas


@needs_ref
@pytest.mark.parametrize("name", NAMES)
def test_port_vs_reference_rankk(name):
    """xBLAT3-style sweep (dims 0 1 2 3 5 9 + two blocked sizes, ld = dim+1, both triangles, every legal trans, the
    .dat alpha/beta grids): the port and the reference agree on the referenced triangle to a gauge ratio of 8, half the netlib threshold (the
    reference sums small n through its scalar-tail micro-kernels, the port with plain per-kc FMA chains), both
    leave the other triangle and the ld padding bit-identical, and both pass the netlib criterion (ratio < 16)."""
    oa.sync_cache_sizes()
    t = name[0]
    RB = oa.ref_blas()
    rng = np.random.default_rng(17)
    transes, alphas, betas = _params(name)
    for n, k in itertools.product((0, 1, 2, 3, 5, 9, 70), (0, 1, 2, 5, 9, 400)):
        for uplo, trans in itertools.product("UL", transes):
            ra, ca = (n, k) if trans == "N" else (k, n)
            A = oa.rand_matrix(rng, t, ra, ca, ld=ra + 1)
            for alpha, beta in itertools.product(alphas, betas):
                C0 = oa.rand_matrix(rng, t, n, n, ld=n + 1)
                c1, c2 = C0.copy(order="F"), C0.copy(order="F")
                assert oa.call_rankk(getattr(RB, name), name, uplo, trans, n, k, alpha, A, ra + 1, beta, c1, n + 1) == 0
                assert oa.call_rankk(getattr(P, "oracle_" + name), name, uplo, trans, n, k, alpha, A, ra + 1, beta, c2, n + 1) == 0
                if n == 0:
                    continue
                m = oa.tri_mask(n, uplo)
                for cc in (c1, c2):
                    assert cc[n:].tobytes() == C0[n:].tobytes()                     # ld padding untouched
                    assert np.array_equal(cc[:n][~m], C0[:n][~m])                    # other triangle untouched
                ref, g = oa.hp_rankk(name, uplo, trans, n, k, alpha, A, ra + 1, beta, C0, n + 1)
                if "herk" in name:
                    # Hermitian diagonal: imaginary part zero whenever the routine writes the diagonal
                    wrote = not (beta == 1.0 and (k == 0 or alpha == 0.0))
                    if wrote:
                        assert np.all(np.diagonal(c1[:n]).imag == 0) and np.all(np.diagonal(c2[:n]).imag == 0)
                        ref[np.arange(n), np.arange(n)] = ref[np.arange(n), np.arange(n)].real
                eps = oa.EPS[t]
                gg = np.maximum(g, 1e-300)
                assert (np.abs(c1[:n] - c2[:n])[m] / (eps * gg[m])).max() < 8.0, (name, n, k, uplo, trans, alpha, beta)
                assert (np.abs(c2[:n] - ref)[m] / (eps * gg[m])).max() < 16.0


@pytest.mark.parametrize("name", NAMES)
def test_port_rankk_error_exits(name):
    """DCHKE's ?SYRK / ?HERK blocks (blas/testing/dblat3.f:2232-2296, zblat3.f): info 1, 2, 3, 4, 7, 10."""
    t = name[0]
    herk, cplx = "herk" in name, t in "cz"
    a = np.zeros((4, 4), dtype=oa.NP_DTYPE[t], order="F")
    c = np.zeros((4, 4), dtype=oa.NP_DTYPE[t], order="F")
    bad_trans = "T" if herk else ("C" if cplx else "/")
    good_t = "C" if herk else "T"
    cases = [(1, "/", "N", 0, 0, 1, 1), (2, "U", bad_trans, 0, 0, 1, 1), (2, "U", "/", 0, 0, 1, 1),
             (3, "U", "N", -1, 0, 1, 1), (3, "L", good_t, -1, 0, 1, 1), (4, "U", "N", 0, -1, 1, 1), (4, "L", good_t, 0, -1, 1, 1),
             (7, "U", "N", 2, 0, 1, 2), (7, "U", good_t, 0, 2, 1, 1), (7, "L", "N", 2, 0, 1, 2), (7, "L", good_t, 0, 2, 1, 1),
             (10, "U", "N", 2, 0, 2, 1), (10, "U", good_t, 2, 0, 1, 1), (10, "L", "N", 2, 0, 2, 1), (10, "L", good_t, 2, 0, 1, 1)]
    label = (name[:-1].upper() + " ").encode()
    targets = [getattr(P, "oracle_" + name)]
    if oa.have_ref():
        targets.append(getattr(oa.ref_blas(), name))
    for fn in targets:
        for (info, uplo, trans, n, k, lda, ldc) in cases:
            P.oracle_xerbla_expect(label, info)
            oa.call_rankk(fn, name, uplo, trans, n, k, 1.0, a, lda, 1.0, c, ldc)
            assert P.oracle_xerbla_result() == 1, (name, info, uplo, trans, n, k, lda, ldc)
            assert not c.any()
