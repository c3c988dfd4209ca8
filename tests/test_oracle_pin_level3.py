"""Pin of oracle/level3_port.c (?trsm_ ?trmm_ ?symm_ ?hemm_ ?syr2k_ ?her2k_) against the reference compiled from
/root/reference (oracle/_ref/libeigen_blas_ref.so) and against the long-double evaluation.  CPU only."""
import itertools

import numpy as np
import pytest

import level3_cases as lc
import oracle_api as oa

P = oa.port()
needs_ref = pytest.mark.skipif(not oa.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
ALL = oa.TRI_NAMES + oa.SYMM_NAMES + oa.R2K_NAMES


def _fns(name):
    fns = [("port", getattr(P, "oracle_" + name))]
    if oa.have_ref():
        fns.append(("ref", getattr(oa.ref_blas(), name)))
    return fns


@pytest.mark.parametrize("name", oa.TRI_NAMES)
def test_tri_sweep_port_and_reference(name):
    for label, fn in _fns(name):
        lc.sweep_tri(fn, name, np.random.default_rng(5), extra=((35, 7), (7, 35)))


@pytest.mark.parametrize("name", oa.SYMM_NAMES)
def test_symm_sweep_port_and_reference(name):
    for label, fn in _fns(name):
        lc.sweep_symm(fn, name, np.random.default_rng(6), extra=((35, 7), (7, 35)))


@pytest.mark.parametrize("name", oa.R2K_NAMES)
def test_r2k_sweep_port_and_reference(name):
    for label, fn in _fns(name):
        lc.sweep_r2k(fn, name, np.random.default_rng(7), extra=((35, 7), (7, 35)))


@needs_ref
@pytest.mark.parametrize("name", ALL)
def test_port_matches_reference_elementwise(name):
    """Same inputs through the port and the compiled reference: results agree to a few eps of the result scale, return
    values are identical."""
    t = name[0]
    rng = np.random.default_rng(11)
    R = oa.ref_blas()
    alphas, betas = lc.scalars(name)
    for (d1, d2) in [(5, 9), (9, 5), (33, 17), (0, 3), (3, 0)]:
        if name[1:] in ("trsm_", "trmm_"):
            for side, uplo, trans, diag in itertools.product("LR", "UL", "NTC", "UN"):
                a, b0 = lc.tri_inputs(rng, name, side, d1, d2)
                bp, br = b0.copy(order="F"), b0.copy(order="F")
                rp = oa.call_tri(getattr(P, "oracle_" + name), name, side, uplo, trans, diag, d1, d2, alphas[2], a, a.shape[0], bp, bp.shape[0])
                rr = oa.call_tri(getattr(R, name), name, side, uplo, trans, diag, d1, d2, alphas[2], a, a.shape[0], br, br.shape[0])
                assert rp == rr
                scale = max(1.0, float(np.abs(br).max())) if br.size else 1.0
                assert np.abs(bp - br).max(initial=0.0) <= 64 * oa.EPS[t] * scale
        elif name[1:] in ("symm_", "hemm_"):
            for side, uplo in itertools.product("LR", "UL"):
                na = d1 if side == "L" else d2
                a = oa.rand_matrix(rng, t, na, na, ld=na + 1)
                b = oa.rand_matrix(rng, t, d1, d2, ld=d1 + 1)
                c0 = oa.rand_matrix(rng, t, d1, d2, ld=d1 + 1)
                cp, cr = c0.copy(order="F"), c0.copy(order="F")
                rp = oa.call_abc(getattr(P, "oracle_" + name), name, side, uplo, d1, d2, alphas[2], a, na + 1, b, d1 + 1, betas[2], cp, d1 + 1)
                rr = oa.call_abc(getattr(R, name), name, side, uplo, d1, d2, alphas[2], a, na + 1, b, d1 + 1, betas[2], cr, d1 + 1)
                assert rp == rr
                assert np.abs(cp - cr).max(initial=0.0) <= 64 * oa.EPS[t] * max(1.0, float(np.abs(cr).max(initial=0.0)))
        else:
            for uplo, trans in itertools.product("UL", lc.legal_trans(name)):
                n, k = d1, d2
                ra, ca = (n, k) if trans == "N" else (k, n)
                a = oa.rand_matrix(rng, t, ra, ca, ld=ra + 1)
                b = oa.rand_matrix(rng, t, ra, ca, ld=ra + 1)
                c0 = oa.rand_matrix(rng, t, n, n, ld=n + 1)
                cp, cr = c0.copy(order="F"), c0.copy(order="F")
                rp = oa.call_abc(getattr(P, "oracle_" + name), name, uplo, trans, n, k, alphas[2], a, ra + 1, b, ra + 1, betas[2], cp, n + 1)
                rr = oa.call_abc(getattr(R, name), name, uplo, trans, n, k, alphas[2], a, ra + 1, b, ra + 1, betas[2], cr, n + 1)
                assert rp == rr
                assert np.abs(cp - cr).max(initial=0.0) <= 64 * oa.EPS[t] * max(1.0, float(np.abs(cr).max(initial=0.0)))


@pytest.mark.parametrize("name", ALL)
def test_error_exits_port_and_reference(name):
    for label, fn in _fns(name):
        lc.run_error_exits(P, fn, name)
