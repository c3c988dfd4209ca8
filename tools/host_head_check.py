#!/usr/bin/env python
"""Host-operand products large enough for the multi-slab / chunked-A head of host.cu::run_host (k >= 4096, n >= 1024), checked on
sampled rows against the long-double oracle: N/T/C operands, alpha / beta != 1, beta = 0 with NaN in C, ld = dim + 1."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import eigen_b200  # noqa: E402
import oracle_api as oa  # noqa: E402

L = eigen_b200.require_device()
rng = np.random.default_rng(5)
for t, ta, tb, m, n, k, al, be in (("d", "N", "N", 3000, 3100, 4500, 0.7, 1.3), ("d", "T", "C", 2100, 2600, 5000, -1.0, 0.0),
                                   ("s", "N", "T", 2500, 4100, 4200, 1.0, 1.0), ("z", "C", "N", 1500, 2100, 4100, 0.7 - 0.9j, 1.3 - 1.1j)):
    ra, ca = (m, k) if ta == "N" else (k, m)
    rb, cb = (k, n) if tb == "N" else (n, k)
    A = oa.rand_matrix(rng, t, ra, ca, ld=ra + 1)
    B = oa.rand_matrix(rng, t, rb, cb, ld=rb + 1)
    C0 = oa.rand_matrix(rng, t, m, n, ld=m + 1)
    if be == 0.0:
        C0[:m] = np.nan
    c = C0.copy(order="F")
    assert oa.call_gemm(getattr(L, t + "gemm_"), t, ta, tb, m, n, k, al, A, ra + 1, B, rb + 1, be, c, m + 1) == 0, eigen_b200.last_error()
    assert c[m:].tobytes() == C0[m:].tobytes(), "padding row touched"
    rows = np.array([0, 1, m // 3, m // 2, m - 2, m - 1], dtype=np.int32)
    Cin = C0 if be != 0.0 else np.zeros_like(C0)
    ref, g = oa.hp_gemm(t, ta, tb, m, n, k, al, A, ra + 1, B, rb + 1, be, Cin, m + 1, rows=rows)
    ratio = float((np.abs(c[rows] - ref) / (oa.EPS[t] * np.maximum(g, 1e-300))).max())
    assert ratio < 16.0, (t, ta, tb, ratio)
    print("ok %sgemm %s%s %dx%dx%d ratio=%.2f variant=%s" % (t, ta, tb, m, n, k, ratio, eigen_b200.last_variant()), flush=True)
print("host_head_check: all cases passed")
