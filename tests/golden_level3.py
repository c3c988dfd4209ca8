"""Replay of the committed reference outputs tests/golden/level3/*.npz (made by tests/golden/make_golden_level3.py from
the reference's own blas/ and lapack/ libraries) through any implementation of the same F77 entry points.
TEST INFRASTRUCTURE: used by tests/test_golden_level3.py (oracle port, CPU) and tests/test_gpu_zz_golden_level3.py (library, GPU)."""
import glob
import os

import numpy as np

import oracle_api as oa

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = sorted(glob.glob(os.path.join(HERE, "golden", "level3", "*.npz")))


def _scalar(x):
    x = np.asarray(x).reshape(-1)[0]
    return complex(x) if np.iscomplexobj(x) else float(x)


def replay(fn_of, path):
    """fn_of(name) -> ctypes function.  Runs the fixture and compares with the reference output: identical pivots and
    info, values within 16 (k + 8) eps of the result scale (k = length of the inner sums; BASELINE.json's tolerance is
    c * k * eps), untouched regions bit-identical to the reference's."""
    z = np.load(path)
    name = str(z["routine"])
    t = name[0]
    eps = oa.EPS[t]
    fn = fn_of(name)
    dims = [int(v) for v in z["dims"]]
    if name[1:] in ("syrk_", "herk_", "syr2k_", "her2k_"):
        n, k = dims
        uplo, trans = str(z["c1"]), str(z["c2"])
        A, B, C0, ref = z["A"], z["B"], z["C0"], z["Cref"]
        c = np.array(C0, order="F")
        if "2k" in name:
            oa.call_abc(fn, name, uplo, trans, n, k, _scalar(z["alpha"]), A, A.shape[0], B, B.shape[0], _scalar(z["beta"]), c, c.shape[0])
        else:
            oa.call_rankk(fn, name, uplo, trans, n, k, _scalar(z["alpha"]), A, A.shape[0], _scalar(z["beta"]), c, c.shape[0])
        mask = np.zeros(c.shape, dtype=bool)
        mask[:n] = oa.tri_mask(n, uplo)
        assert c[~mask].tobytes() == ref[~mask].tobytes(), "the other triangle / padding differs from the reference"
        inner = (2 if "2k" in name else 1) * k
        got, want = c[mask], ref[mask]
    elif name[1:] in ("symm_", "hemm_"):
        m, n = dims
        side, uplo = str(z["c1"]), str(z["c2"])
        A, B, C0, ref = z["A"], z["B"], z["C0"], z["Cref"]
        c = np.array(C0, order="F")
        oa.call_abc(fn, name, side, uplo, m, n, _scalar(z["alpha"]), A, A.shape[0], B, B.shape[0], _scalar(z["beta"]), c, c.shape[0])
        assert c[m:].tobytes() == ref[m:].tobytes()
        inner = m if side == "L" else n
        got, want = c[:m], ref[:m]
    elif name[1:] in ("trsm_", "trmm_"):
        m, n = dims
        A, B0, ref = z["A"], z["B0"], z["Bref"]
        b = np.array(B0, order="F")
        oa.call_tri(fn, name, str(z["c1"]), str(z["c2"]), str(z["c3"]), str(z["c4"]), m, n, _scalar(z["alpha"]), A, A.shape[0], b, b.shape[0])
        assert b[m:].tobytes() == ref[m:].tobytes()
        inner = m if str(z["c1"]) == "L" else n
        got, want = b[:m], ref[:m]
    elif name[1:] == "potrf_":
        (n,) = dims
        uplo = str(z["c1"])
        a = np.array(z["A0"], order="F")
        ref = z["Aref"]
        info = oa.call_potrf(fn, uplo, n, a, a.shape[0])
        assert info == int(z["info"])
        mask = np.zeros(a.shape, dtype=bool)
        mask[:n] = oa.tri_mask(n, uplo)
        assert a[~mask].tobytes() == ref[~mask].tobytes(), "the other triangle / padding differs from the reference"
        inner = n
        got, want = a[mask], ref[mask]
    else:
        m, n = dims
        a = np.array(z["A0"], order="F")
        ref = z["Aref"]
        ipiv, info = oa.call_getrf(fn, m, n, a, a.shape[0])
        assert info == int(z["info"])
        assert np.array_equal(ipiv, z["ipiv"]), "pivot sequence differs from the reference"
        assert a[m:].tobytes() == ref[m:].tobytes()
        inner = min(m, n)
        got, want = a[:m], ref[:m]
    scale = max(1.0, float(np.abs(want).max(initial=0.0)))
    err = float(np.abs(got - want).max(initial=0.0))
    assert err <= 16.0 * (inner + 8) * eps * scale, (os.path.basename(path), err, 16.0 * (inner + 8) * eps * scale)
    return err
