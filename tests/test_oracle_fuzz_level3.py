"""Seeded random-shape fuzz of oracle/level3_port.c, rankk_port.c and lapack_port.c against the reference compiled from
/root/reference: ragged dimensions, random flags, random leading-dimension padding.  CPU only; skipped without oracle/_ref."""
import numpy as np
import pytest

import lapack_cases as lp
import level3_cases as lc
import oracle_api as oa

P = oa.port()
pytestmark = pytest.mark.skipif(not (oa.have_ref() and oa.have_ref_lapack()), reason="oracle/_ref not built (needs /root/reference)")


def _close(t, a, b, k):
    scale = max(1.0, float(np.abs(b).max(initial=0.0)))
    with np.errstate(invalid="ignore"):
        d = np.abs(a - b)
    d = np.where(np.isnan(a) & np.isnan(b), 0.0, d)
    return float(d.max(initial=0.0)) <= 16.0 * (k + 8) * oa.EPS[t] * scale


@pytest.mark.parametrize("seed", range(6))
def test_blas3_fuzz(seed):
    rng = np.random.default_rng(1000 + seed)
    R = oa.ref_blas()
    for _ in range(40):
        name = str(rng.choice(oa.RANKK_NAMES + oa.R2K_NAMES + oa.SYMM_NAMES + oa.TRI_NAMES))
        t = name[0]
        d1, d2 = int(rng.integers(0, 48)), int(rng.integers(0, 48))
        pad = int(rng.integers(0, 3))
        alphas, betas = lc.scalars(name) if name not in oa.RANKK_NAMES else (([0.0, 1.0, 0.7], [0.0, 1.0, 1.3]) if ("herk" in name or t in "sd") else ([0.0, 1.0, 0.7 - 0.9j], [0.0, 1.0, 1.3 - 1.1j]))
        alpha, beta = alphas[int(rng.integers(0, 3))], betas[int(rng.integers(0, 3))]
        uplo, side = str(rng.choice(list("UL"))), str(rng.choice(list("LR")))
        if name in oa.TRI_NAMES:
            trans, diag = str(rng.choice(list("NTC"))), str(rng.choice(list("UN")))
            a, b0 = lc.tri_inputs(rng, name, side, d1, d2, lda_pad=pad, ldb_pad=pad)
            bp, br = b0.copy(order="F"), b0.copy(order="F")
            rp = oa.call_tri(getattr(P, "oracle_" + name), name, side, uplo, trans, diag, d1, d2, alpha, a, a.shape[0], bp, bp.shape[0])
            rr = oa.call_tri(getattr(R, name), name, side, uplo, trans, diag, d1, d2, alpha, a, a.shape[0], br, br.shape[0])
            assert rp == rr and _close(t, bp, br, max(d1, d2)), (name, side, uplo, trans, diag, d1, d2, alpha)
        elif name in oa.SYMM_NAMES:
            na = d1 if side == "L" else d2
            a = oa.rand_matrix(rng, t, na, na, ld=na + pad)
            b = oa.rand_matrix(rng, t, d1, d2, ld=d1 + pad)
            c0 = oa.rand_matrix(rng, t, d1, d2, ld=d1 + pad)
            cp, cr = c0.copy(order="F"), c0.copy(order="F")
            rp = oa.call_abc(getattr(P, "oracle_" + name), name, side, uplo, d1, d2, alpha, a, a.shape[0], b, b.shape[0], beta, cp, cp.shape[0])
            rr = oa.call_abc(getattr(R, name), name, side, uplo, d1, d2, alpha, a, a.shape[0], b, b.shape[0], beta, cr, cr.shape[0])
            assert rp == rr and _close(t, cp, cr, na), (name, side, uplo, d1, d2, alpha, beta)
        else:
            n, k = d1, d2
            two = name in oa.R2K_NAMES
            legal = lc.legal_trans(name) if two else ("NC" if "herk" in name else ("NT" if t in "cz" else "NTC"))
            trans = str(rng.choice(list(legal)))
            ra, ca = (n, k) if trans == "N" else (k, n)
            a = oa.rand_matrix(rng, t, ra, ca, ld=ra + pad)
            b = oa.rand_matrix(rng, t, ra, ca, ld=ra + pad)
            c0 = oa.rand_matrix(rng, t, n, n, ld=n + pad)
            cp, cr = c0.copy(order="F"), c0.copy(order="F")
            if two:
                rp = oa.call_abc(getattr(P, "oracle_" + name), name, uplo, trans, n, k, alpha, a, a.shape[0], b, b.shape[0], beta, cp, cp.shape[0])
                rr = oa.call_abc(getattr(R, name), name, uplo, trans, n, k, alpha, a, a.shape[0], b, b.shape[0], beta, cr, cr.shape[0])
            else:
                rp = oa.call_rankk(getattr(P, "oracle_" + name), name, uplo, trans, n, k, alpha, a, a.shape[0], beta, cp, cp.shape[0])
                rr = oa.call_rankk(getattr(R, name), name, uplo, trans, n, k, alpha, a, a.shape[0], beta, cr, cr.shape[0])
            assert rp == rr and _close(t, cp, cr, 2 * k), (name, uplo, trans, n, k, alpha, beta)


@pytest.mark.parametrize("seed", range(3))
def test_lapack_fuzz(seed):
    rng = np.random.default_rng(2000 + seed)
    R = oa.ref_lapack()
    for _ in range(20):
        t = str(rng.choice(list("sdcz")))
        n = int(rng.integers(0, 90))
        uplo = str(rng.choice(list("UL")))
        full = lp.make_hpd(rng, t, n, ld=n + int(rng.integers(0, 3)))
        ap, ar = full.copy(order="F"), full.copy(order="F")
        ip_, ir = oa.call_potrf(getattr(P, "oracle_%spotrf_" % t), uplo, n, ap, ap.shape[0]), oa.call_potrf(getattr(R, t + "potrf_"), uplo, n, ar, ar.shape[0])
        assert ip_ == ir and _close(t, ap, ar, n), (t, "potrf", uplo, n)
        m = n + int(rng.integers(0, 40))   # m >= n: the shapes the reference's blocked_lu completes
        a0 = oa.rand_matrix(rng, t, m, n, ld=m + int(rng.integers(0, 3)))
        ap, ar = a0.copy(order="F"), a0.copy(order="F")
        pp, ip_ = oa.call_getrf(getattr(P, "oracle_%sgetrf_" % t), m, n, ap, ap.shape[0])
        pr, ir = oa.call_getrf(getattr(R, t + "getrf_"), m, n, ar, ar.shape[0])
        assert ip_ == ir and np.array_equal(pp, pr) and _close(t, ap, ar, 16 * max(n, 1)), (t, "getrf", m, n)


def test_reference_getrf_leaves_wide_matrices_unfinished_and_the_port_reproduces_it():
    """for m < n (min(m, n) > 16) the reference's blocked_lu never updates the columns right of the
    square part (PartialPivLU.h:457-492 uses tsize = size - k - bs), so its output does not satisfy P A = L U.  The
    oracle port restates that behaviour bit for pivot; the GPU library deliberately completes the factorization (LAPACK
    semantics) and parity with the reference is claimed for m >= n only."""
    rng = np.random.default_rng(5)
    R = oa.ref_lapack()
    for (m, n) in [(20, 37), (33, 70)]:
        a0 = oa.rand_matrix(rng, "d", m, n)
        ap, ar = a0.copy(order="F"), a0.copy(order="F")
        pp, ip_ = oa.call_getrf(P.oracle_dgetrf_, m, n, ap, m)
        pr, ir = oa.call_getrf(R.dgetrf_, m, n, ar, m)
        assert ip_ == ir and np.array_equal(pp, pr) and _close("d", ap, ar, 16 * m)
        with pytest.raises(AssertionError):
            lp.check_getrf("d", m, n, a0, ar, pr, ir)
