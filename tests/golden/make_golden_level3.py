"""Generates tests/golden/level3/*.npz from the REFERENCE ITSELF: oracle/_ref/libeigen_blas_ref.so (blas/level3_impl.h
syrk, herk, syr2k, her2k, symm, hemm, trsm, trmm) and oracle/_ref/libeigen_lapack_ref.so (lapack/cholesky.cpp potrf,
lapack/lu.cpp getrf), both compiled unmodified from /root/reference by oracle/Makefile.  Run in the build container:

    python tests/golden/make_golden_level3.py

Each fixture stores the routine name, its character / integer arguments, the input operands and what the reference left
in the output operand (plus ipiv / info for the factorizations)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import lapack_cases as lp  # noqa: E402
import level3_cases as lc  # noqa: E402
import oracle_api as oa  # noqa: E402

OUT = os.path.join(HERE, "level3")
AL = {"s": 0.7, "d": 0.7, "c": 0.7 - 0.9j, "z": 0.7 - 0.9j}
BE = {"s": 1.3, "d": 1.3, "c": 1.3 - 1.1j, "z": 1.3 - 1.1j}


def save(name, **kw):
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **kw)
    print("wrote", name)


def main():
    os.makedirs(OUT, exist_ok=True)
    RB, RL = oa.ref_blas(), oa.ref_lapack()
    rng = np.random.default_rng(20261018)
    idx = 0
    # rank-k and rank-2k updates
    for name, uplo, trans, n, k in [("dsyrk_", "L", "N", 70, 129), ("ssyrk_", "U", "T", 45, 300), ("csyrk_", "U", "N", 33, 64),
                                    ("zherk_", "L", "C", 40, 77), ("cherk_", "U", "N", 52, 31),
                                    ("dsyr2k_", "U", "T", 61, 90), ("zsyr2k_", "L", "N", 30, 45), ("zher2k_", "U", "C", 36, 58),
                                    ("cher2k_", "L", "N", 44, 37)]:
        t = name[0]
        herk, her2k, two = "herk" in name, "her2k" in name, "2k" in name
        ra, ca = (n, k) if trans == "N" else (k, n)
        A = oa.rand_matrix(rng, t, ra, ca, ld=ra + 1)
        B = oa.rand_matrix(rng, t, ra, ca, ld=ra + 2)
        C0 = oa.rand_matrix(rng, t, n, n, ld=n + 1)
        alpha = 0.7 if herk else AL[t]
        beta = 1.3 if (herk or her2k) else BE[t]
        Cref = C0.copy(order="F")
        if two:
            oa.call_abc(getattr(RB, name), name, uplo, trans, n, k, alpha, A, ra + 1, B, ra + 2, beta, Cref, n + 1)
        else:
            oa.call_rankk(getattr(RB, name), name, uplo, trans, n, k, alpha, A, ra + 1, beta, Cref, n + 1)
        save("%02d_%s%s%s_%dx%d" % (idx, name, uplo, trans, n, k), routine=name, c1=uplo, c2=trans, dims=[n, k], alpha=np.array(alpha),
             beta=np.array(beta), A=A, B=B, C0=C0, Cref=Cref)
        idx += 1
    # symmetric / Hermitian products
    for name, side, uplo, m, n in [("dsymm_", "L", "U", 70, 41), ("ssymm_", "R", "L", 38, 90), ("zhemm_", "L", "L", 45, 33), ("chemm_", "R", "U", 29, 64),
                                   ("zsymm_", "R", "U", 31, 40)]:
        t = name[0]
        na = m if side == "L" else n
        A = oa.rand_matrix(rng, t, na, na, ld=na + 1)
        B = oa.rand_matrix(rng, t, m, n, ld=m + 2)
        C0 = oa.rand_matrix(rng, t, m, n, ld=m + 1)
        Cref = C0.copy(order="F")
        oa.call_abc(getattr(RB, name), name, side, uplo, m, n, AL[t], A, na + 1, B, m + 2, BE[t], Cref, m + 1)
        save("%02d_%s%s%s_%dx%d" % (idx, name, side, uplo, m, n), routine=name, c1=side, c2=uplo, dims=[m, n], alpha=np.array(AL[t]),
             beta=np.array(BE[t]), A=A, B=B, C0=C0, Cref=Cref)
        idx += 1
    # triangular solves and products
    for name, side, uplo, trans, diag, m, n in [("dtrsm_", "L", "L", "N", "U", 150, 60), ("dtrsm_", "R", "L", "C", "N", 90, 140), ("strsm_", "L", "U", "T", "N", 64, 77),
                                                ("ztrsm_", "R", "U", "N", "N", 40, 70), ("ctrsm_", "L", "L", "C", "U", 55, 30),
                                                ("dtrmm_", "L", "U", "N", "N", 130, 50), ("strmm_", "R", "L", "T", "U", 60, 140), ("ztrmm_", "L", "L", "C", "N", 70, 20),
                                                ("ctrmm_", "R", "U", "N", "U", 33, 45)]:
        t = name[0]
        A, B0 = lc.tri_inputs(rng, name, side, m, n)
        Bref = B0.copy(order="F")
        oa.call_tri(getattr(RB, name), name, side, uplo, trans, diag, m, n, AL[t], A, A.shape[0], Bref, Bref.shape[0])
        save("%02d_%s%s%s%s%s_%dx%d" % (idx, name, side, uplo, trans, diag, m, n), routine=name, c1=side, c2=uplo, c3=trans, c4=diag, dims=[m, n],
             alpha=np.array(AL[t]), A=A, B0=B0, Bref=Bref)
        idx += 1
    # factorizations
    for t, uplo, n in [("d", "L", 100), ("d", "U", 67), ("s", "L", 90), ("z", "U", 50), ("c", "L", 41)]:
        A0 = lp.make_hpd(rng, t, n, ld=n + 1)
        Aref = A0.copy(order="F")
        info = oa.call_potrf(getattr(RL, t + "potrf_"), uplo, n, Aref, n + 1)
        save("%02d_%spotrf_%s_%d" % (idx, t, uplo, n), routine=t + "potrf_", c1=uplo, dims=[n], A0=A0, Aref=Aref, info=info)
        idx += 1
    for t, m, n in [("d", 100, 100), ("d", 130, 60), ("s", 90, 90), ("z", 50, 50), ("c", 64, 33)]:
        A0 = oa.rand_matrix(rng, t, m, n, ld=m + 1)
        Aref = A0.copy(order="F")
        ipiv, info = oa.call_getrf(getattr(RL, t + "getrf_"), m, n, Aref, m + 1)
        save("%02d_%sgetrf_%dx%d" % (idx, t, m, n), routine=t + "getrf_", dims=[m, n], A0=A0, Aref=Aref, ipiv=ipiv, info=info)
        idx += 1


if __name__ == "__main__":
    main()
