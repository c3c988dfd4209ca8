"""libb200blas.so against the committed outputs of the reference's own blas/ and lapack/ libraries -- needs a B200."""
import os

import pytest

import eigen_b200
import golden_level3 as gl

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("path", gl.FIXTURES, ids=[os.path.basename(p)[:-4] for p in gl.FIXTURES])
def test_library_reproduces_reference_outputs(path):
    L = eigen_b200.require_device()
    gl.replay(lambda name: getattr(L, name), path)


@pytest.mark.parametrize("t", list("sdcz"))
def test_getrf_pivot_ties_take_the_first_row(t):
    """maxCoeff keeps the FIRST largest entry (PartialPivLU.h:378-380): a first column of equal magnitudes pivots on row 1,
    whatever CTA of the cooperative panel kernel owns the other candidates; an all-zero first column reports info = 1."""
    import numpy as np
    import oracle_api as oa
    L = eigen_b200.require_device()
    rng = np.random.default_rng(12)
    for m, n in ((700, 40), (3000, 33)):   # several panel CTAs
        a = oa.rand_matrix(rng, t, m, n)
        a[:, 0] = np.where(np.arange(m) % 2 == 0, 1.0, -1.0)
        ipiv, info = oa.call_getrf(getattr(L, t + "getrf_"), m, n, a.copy(order="F"), m)
        assert info == 0 and ipiv[0] == 1, (t, m, n, ipiv[:4], info)
        a[:, 0] = 0
        ipiv, info = oa.call_getrf(getattr(L, t + "getrf_"), m, n, a.copy(order="F"), m)
        assert info == 1 and ipiv[0] == 1, (t, m, n, ipiv[:4], info)
