#!/usr/bin/env python
"""Small instances of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool racecheck python tools/sanitize_small.py

Each case is checked against the long-double oracle as well, so a sanitizer-clean run is also a correct one.  Sizes are
chosen so that every pipeline runs a few stages (TMA ring wrap-around, both TMEM buffers, CTA-pair kernels, the mbarrier
DMMA ring, the cluster LU panel, the one-CTA Cholesky block) while the whole script stays within minutes under the tool."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import eigen_b200  # noqa: E402
import oracle_api as oa  # noqa: E402
import lapack_cases as lp  # noqa: E402
import level3_cases as lc  # noqa: E402

L = eigen_b200.require_device()
rng = np.random.default_rng(3)


def gemm_case(t, ta, tb, m, n, k, env=None):
    for key, val in (env or {}).items():
        os.environ[key] = val
    A = oa.rand_matrix(rng, t, *( (m, k) if ta == "N" else (k, m)))
    B = oa.rand_matrix(rng, t, *( (k, n) if tb == "N" else (n, k)))
    C0 = oa.rand_matrix(rng, t, m, n)
    c = C0.copy(order="F")
    r = eigen_b200.gemm_host(t, ta, tb, m, n, k, 0.7, A, A.shape[0], B, B.shape[0], 1.3, c, m)
    assert r == 0, eigen_b200.last_error()
    ref, g = oa.hp_gemm(t, ta, tb, m, n, k, 0.7, A, A.shape[0], B, B.shape[0], 1.3, C0, m)
    ratio = float((np.abs(c - ref) / (oa.EPS[t] * g)).max())
    assert ratio < 16.0, (t, ta, tb, m, n, k, ratio)
    print("ok %sgemm %s%s %dx%dx%d variant=%s ratio=%.2f" % (t, ta, tb, m, n, k, eigen_b200.last_variant(), ratio), flush=True)


# DMMA: big tile (mbarrier ring wraps: k = 200 > 4 stages x 16), small tile, transposed loaders, complex
gemm_case("d", "N", "N", 300, 200, 200)
gemm_case("d", "T", "C", 1400, 1300, 100)
gemm_case("z", "C", "N", 150, 130, 90)
# tcgen05 3xTF32: 1-CTA kernel (TMA ring of 2 wraps, both TMEM buffers: k = 300 > kchunk x 32), complex, then the CTA-pair kernels
gemm_case("s", "N", "T", 300, 280, 300)
gemm_case("c", "N", "C", 200, 150, 140)
os.environ["B200BLAS_TF32_PAIR"] = "1"
gemm_case("s", "T", "N", 700, 600, 200)
gemm_case("c", "N", "N", 600, 520, 100)
os.environ.pop("B200BLAS_TF32_PAIR")
# SIMT
gemm_case("s", "N", "N", 40, 30, 20)
gemm_case("z", "T", "N", 33, 17, 9)
# triangular solve (substitution leaf + product), Cholesky (one-CTA block + leaf + rank-k), LU (cluster register panel)
Lb = eigen_b200.lib()
a, b0 = lc.tri_inputs(rng, "dtrsm_", "L", 300, 200)
b = b0.copy(order="F")
oa.call_tri(Lb.dtrsm_, "dtrsm_", "L", "L", "N", "N", 300, 200, 0.7, a, a.shape[0], b, b.shape[0])
print("ok dtrsm ratio=%.2f" % lc.check_tri("dtrsm_", "L", "L", "N", "N", 300, 200, 0.7, a, b0, b), flush=True)
for t, n in (("d", 700), ("z", 200), ("s", 300)):
    for uplo in "LU":
        full = lp.make_hpd(rng, t, n)
        f = full.copy(order="F")
        info = oa.call_potrf(getattr(Lb, t + "potrf_"), uplo, n, f, n)
        print("ok %spotrf %s %d ratio=%.2f" % (t, uplo, n, lp.check_potrf(t, uplo, n, full, full, f, info)), flush=True)
for t, m, n in (("d", 1300, 700), ("c", 300, 300), ("s", 700, 900)):
    g0 = oa.rand_matrix(rng, t, m, n)
    lu = g0.copy(order="F")
    ipiv, info = oa.call_getrf(getattr(Lb, t + "getrf_"), m, n, lu, m)
    print("ok %sgetrf %dx%d ratio=%.2f" % (t, m, n, lp.check_getrf(t, m, n, g0, lu, ipiv, info)), flush=True)
print("sanitize_small: all cases passed")
