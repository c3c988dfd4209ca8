"""Generates tests/golden/*.npz from the REFERENCE ITSELF (oracle/_ref/libeigen_blas_ref.so, i.e. the unmodified
blas/level3_impl.h gemm compiled from /root/reference by oracle/Makefile).  Run in the build container:

    python tests/golden/make_golden.py

Each fixture stores the inputs, the reference output, and the cache sizes the reference detected on the
generating host (they determine kc/mc/nc and therefore rounding)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_api as oa  # noqa: E402

CASES = [  # (type, ta, tb, m, n, k, alpha, beta)
    ("d", "N", "N", 96, 80, 400, 1.0, 1.0),
    ("d", "T", "N", 61, 47, 129, -1.0, 1.0),
    ("d", "N", "T", 128, 33, 17, 0.7, 1.3),
    ("s", "N", "N", 100, 90, 500, 1.0, 1.0),
    ("s", "C", "T", 45, 77, 63, 0.7, 0.0),
    ("z", "N", "N", 40, 36, 200, 1.0, 1.0),
    ("z", "C", "N", 33, 21, 64, 0.7 - 0.9j, 1.3 - 1.1j),
    ("z", "N", "C", 18, 40, 31, 0.7 - 0.9j, 0.0),
    ("c", "N", "N", 44, 52, 300, 1.0, 1.0),
    ("c", "T", "C", 27, 35, 50, 0.7 - 0.9j, 1.3 - 1.1j),
]


def main():
    cache = oa.sync_cache_sizes()
    RB = oa.ref_blas()
    rng = np.random.default_rng(20261017)
    for i, (t, ta, tb, m, n, k, al, be) in enumerate(CASES):
        A = oa.rand_matrix(rng, t, m if ta == "N" else k, k if ta == "N" else m, ld=(m if ta == "N" else k) + 1)
        B = oa.rand_matrix(rng, t, k if tb == "N" else n, n if tb == "N" else k, ld=(k if tb == "N" else n) + 2)
        C0 = oa.rand_matrix(rng, t, m, n, ld=m + 1)
        Cref = C0.copy(order="F")
        oa.call_gemm(getattr(RB, t + "gemm_"), t, ta, tb, m, n, k, al, A, A.shape[0], B, B.shape[0], be, Cref, m + 1)
        name = "gemm_%02d_%s_%s%s_%dx%dx%d.npz" % (i, t, ta, tb, m, n, k)
        np.savez_compressed(os.path.join(HERE, name), t=t, ta=ta, tb=tb, mnk=[m, n, k], alpha=np.array(al), beta=np.array(be),
                            A=A, B=B, C0=C0, Cref=Cref, cache=list(cache))
        print("wrote", name)


if __name__ == "__main__":
    main()
