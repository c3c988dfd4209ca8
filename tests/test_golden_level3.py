"""The oracle port against the committed outputs of the reference's own blas/ and lapack/ libraries (CPU)."""
import os

import pytest

import golden_level3 as gl
import oracle_api as oa

P = oa.port()


def test_fixtures_present():
    assert len(gl.FIXTURES) >= 30


@pytest.mark.parametrize("path", gl.FIXTURES, ids=[os.path.basename(p)[:-4] for p in gl.FIXTURES])
def test_oracle_port_reproduces_reference_outputs(path):
    gl.replay(lambda name: getattr(P, "oracle_" + name), path)
