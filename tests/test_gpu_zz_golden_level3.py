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
