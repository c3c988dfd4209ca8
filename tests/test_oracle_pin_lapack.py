"""Pin of oracle/lapack_port.c (?potrf_, ?getrf_) against the reference's own lapack/ routines compiled from
/root/reference (oracle/_ref/libeigen_lapack_ref.so) and against the long-double product of the factors.  CPU only."""
import numpy as np
import pytest

import lapack_cases as lp
import oracle_api as oa

P = oa.port()
needs_ref = pytest.mark.skipif(not oa.have_ref_lapack(), reason="oracle/_ref/libeigen_lapack_ref.so not built")


def _fns(name):
    fns = [("port", getattr(P, "oracle_" + name))]
    if oa.have_ref_lapack():
        fns.append(("ref", getattr(oa.ref_lapack(), name)))
    return fns


@pytest.mark.parametrize("t", list("sdcz"))
def test_potrf_port_and_reference(t):
    for label, fn in _fns(t + "potrf_"):
        rng = np.random.default_rng(3)
        for n in lp.POTRF_SIZES:
            for uplo in "LU":
                full = lp.make_hpd(rng, t, n, ld=n + 1)
                a0 = lp.poison_other_triangle(full.copy(order="F"), n, uplo)
                a = a0.copy(order="F")
                info = oa.call_potrf(fn, uplo, n, a, n + 1)
                lp.check_potrf(t, uplo, n, full, a0, a, info)


@pytest.mark.parametrize("t", list("sdcz"))
def test_getrf_port_and_reference(t):
    for label, fn in _fns(t + "getrf_"):
        rng = np.random.default_rng(4)
        for (m, n) in lp.GETRF_SHAPES:
            a0 = oa.rand_matrix(rng, t, m, n, ld=m + 1)
            a = a0.copy(order="F")
            ipiv, info = oa.call_getrf(fn, m, n, a, m + 1)
            lp.check_getrf(t, m, n, a0, a, ipiv, info)


@needs_ref
@pytest.mark.parametrize("t", list("sdcz"))
def test_port_matches_reference(t):
    """Identical info and pivots; factors agree to a few eps of their scale."""
    R = oa.ref_lapack()
    rng = np.random.default_rng(5)
    for n in (5, 33, 100, 257):
        for uplo in "LU":
            full = lp.make_hpd(rng, t, n)
            ap, ar = full.copy(order="F"), full.copy(order="F")
            ip_, ir = oa.call_potrf(getattr(P, "oracle_%spotrf_" % t), uplo, n, ap, n), oa.call_potrf(getattr(R, t + "potrf_"), uplo, n, ar, n)
            assert ip_ == ir == 0
            assert np.abs(ap - ar).max() <= 256 * oa.EPS[t] * np.abs(ar).max()
    for (m, n) in ((5, 5), (33, 33), (100, 100), (257, 257), (300, 64)):
        a0 = oa.rand_matrix(rng, t, m, n)
        ap, ar = a0.copy(order="F"), a0.copy(order="F")
        pp, ip_ = oa.call_getrf(getattr(P, "oracle_%sgetrf_" % t), m, n, ap, m)
        pr, ir = oa.call_getrf(getattr(R, t + "getrf_"), m, n, ar, m)
        assert ip_ == ir == 0
        assert np.array_equal(pp, pr), "pivot sequences differ"
        assert np.abs(ap - ar).max() <= 4096 * oa.EPS[t] * np.abs(ar).max()


@pytest.mark.parametrize("t", list("sdcz"))
def test_failure_reports_port_and_reference(t):
    """info > 0: first non-positive pivot (potrf, LLT.h:316-317) / first exactly-zero pivot (getrf, PartialPivLU.h:396-401)."""
    rng = np.random.default_rng(6)
    for label, fn in _fns(t + "potrf_"):
        for n, k in ((9, 4), (100, 57), (257, 200)):
            a = lp.make_hpd(rng, t, n)
            a[k, k] = -1.0
            assert oa.call_potrf(fn, "L", n, a.copy(order="F"), n) == k + 1
            assert oa.call_potrf(fn, "U", n, a.copy(order="F"), n) == k + 1
    for label, fn in _fns(t + "getrf_"):
        for (m, n, k) in ((9, 9, 4), (100, 100, 57), (257, 130, 100)):
            a0 = oa.rand_matrix(rng, t, m, n)
            a0[:, k] = 0
            a = a0.copy(order="F")
            ipiv, info = oa.call_getrf(fn, m, n, a, m)
            lp.check_getrf(t, m, n, a0, a, ipiv, info, expect_info=k + 1)


@pytest.mark.parametrize("t", list("sdcz"))
def test_error_exits_port_and_reference(t):
    for label, fn in _fns(t + "potrf_"):
        lp.run_potrf_error_exits(P, fn, t)
    for label, fn in _fns(t + "getrf_"):
        lp.run_getrf_error_exits(P, fn, t)


@pytest.mark.parametrize("t", list("sdcz"))
def test_pivot_ties_take_the_first_row_port_and_reference(t):
    """maxCoeff keeps the first largest entry (PartialPivLU.h:378-380); an all-zero first column reports info = 1."""
    rng = np.random.default_rng(12)
    for label, fn in _fns(t + "getrf_"):
        for m, n in ((70, 40), (300, 33)):
            a = oa.rand_matrix(rng, t, m, n)
            a[:, 0] = np.where(np.arange(m) % 2 == 0, 1.0, -1.0)
            ipiv, info = oa.call_getrf(fn, m, n, a.copy(order="F"), m)
            assert info == 0 and ipiv[0] == 1, (label, t, m, n)
            a[:, 0] = 0
            ipiv, info = oa.call_getrf(fn, m, n, a.copy(order="F"), m)
            assert info == 1 and ipiv[0] == 1, (label, t, m, n)
