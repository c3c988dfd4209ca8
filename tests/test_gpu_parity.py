"""Parity of the sm_100a library against the CPU oracle, through the C ABI (run with -m gpu on a B200).

Bar (BASELINE.json north_star): relative Frobenius error <= c*k*eps of the scalar type against the long-double
reference, here with c = 1 (TOL_FRO), plus the netlib component-wise gauge ratio |c-c_ref|/(eps*G) < 16 that the
reference's own blas/testing harness applies (dblat3.f:2587-2596).  The oracle port's (= Eigen gebp's) own error
against the same reference is computed next to ours and we must stay within 4x of it or under 2*sqrt(k)*eps.
"""
import ctypes as C
import glob
import os

import numpy as np
import pytest

import eigen_b200
import oracle_api as oa

pytestmark = pytest.mark.gpu
P = oa.port()
HERE = os.path.dirname(os.path.abspath(__file__))
TOL_FRO = 1.0     # * k * eps
TOL_RATIO = 16.0  # xBLAT3 THRESH, blas/testing/dblat3.dat:8


@pytest.fixture(scope="module")
def L():
    lib = eigen_b200.require_device()
    yield lib
    lib.b200blas_set_variant(0)


def _variants(t):
    return ["simt", "dmma"] if t in "dz" else ["simt", "tf32x3"]


def _check(t, ta, tb, m, n, k, al, be, A, B, C0, c, ldc, rows=None, port_too=True):
    ref, g = oa.hp_gemm(t, ta, tb, m, n, k, al, A, A.shape[0], B, B.shape[0], be, C0, ldc, rows=rows)
    got = c[:m] if rows is None else c[rows]
    eps = oa.EPS[t]
    ratio = (np.abs(got - ref) / (eps * np.maximum(g, 1e-300))).max() if m and n else 0.0
    fro = np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-300)
    assert ratio < TOL_RATIO, (t, ta, tb, m, n, k, "gauge ratio", ratio, eigen_b200.last_variant())
    assert fro <= TOL_FRO * max(k, 1) * eps, (t, ta, tb, m, n, k, "fro", fro)
    if port_too and rows is None and m * n * k <= 4e8:
        c2 = C0.copy(order="F")
        oa.call_gemm(getattr(P, "oracle_%sgemm_" % t), t, ta, tb, m, n, k, al, A, A.shape[0], B, B.shape[0], be, c2, ldc)
        fro_port = np.linalg.norm(c2[:m] - ref) / max(np.linalg.norm(ref), 1e-300)
        # 3xTF32 drops the lo*lo term and accumulates in (non-IEEE-rounded) fp32 TMEM: allow twice the slack
        slack = 2.0 if t in "sc" else 1.0
        assert fro <= slack * max(4.0 * fro_port, 2.0 * np.sqrt(max(k, 1)) * eps), (t, m, n, k, fro, fro_port, eigen_b200.last_variant())
    return fro, ratio


@pytest.mark.parametrize("t", list("sdcz"))
def test_xblat3_sweep_and_error_exits(L, t):
    """The reference's own acceptance test for ?gemm_ (blas/testing/*blat3.f), restated in oracle/blat3_port.c,
    run against the F77 entry points with host arrays: 17496 calls, ld = dim+1, N/T/C, alpha/beta grids, argument
    preservation incl. the padding rows of C, then the 28 error exits."""
    for v in _variants(t):
        L.b200blas_set_variant(eigen_b200.VARIANT[v])
        rep = oa.Blat3Report()
        P.oracle_blat3_chk1(oa.TYPES[t], C.cast(getattr(L, t + "gemm_"), C.c_void_p), C.byref(rep))
        assert rep.ncalls == 17496
        assert not rep.fatal, (v, rep.msg)
        assert rep.errmax < 16.0, (v, rep.errmax)
    L.b200blas_set_variant(0)
    log = C.create_string_buffer(4096)
    assert P.oracle_blat3_chke(oa.TYPES[t], C.cast(getattr(L, t + "gemm_"), C.c_void_p), None, log, 4096) == 0, log.value


SHAPES = [(1, 1, 1), (5, 3, 2), (17, 9, 33), (64, 64, 64), (65, 63, 67), (128, 128, 128), (129, 257, 40),
          (200, 180, 700), (515, 130, 401), (96, 7, 1000), (7, 300, 129), (256, 512, 1024), (1000, 1, 1000),
          (1, 1000, 1000), (333, 444, 5)]


@pytest.mark.parametrize("t", list("sdcz"))
def test_random_shapes_all_ops_vs_oracle(L, t):
    """Seeded uniform[-1,1] inputs like bench/bench_gemm.cpp:211-214; every op pair, ragged sizes, ld > dim,
    the xBLAT alpha/beta values; each kernel variant forced in turn, then the automatic choice."""
    rng = np.random.default_rng(1234)
    cplx = t in "cz"
    alphas = [1.0, (0.7 - 0.9j) if cplx else 0.7, -1.0]
    betas = [1.0, (1.3 - 1.1j) if cplx else 1.3, 0.0]
    worst = 0.0
    for v in _variants(t) + ["auto"]:
        L.b200blas_set_variant(eigen_b200.VARIANT[v])
        for si, (m, n, k) in enumerate(SHAPES):
            for oi, (ta, tb) in enumerate([(x, y) for x in "NTC" for y in "NTC"]):
                if (si + oi) % 3 and m * n * k > 1e6:
                    continue  # large shapes: a third of the op pairs per shape, rotating
                ra, ca = (m, k) if ta == "N" else (k, m)
                rb, cb = (k, n) if tb == "N" else (n, k)
                A = oa.rand_matrix(rng, t, ra, ca, ld=ra + (si % 3))
                B = oa.rand_matrix(rng, t, rb, cb, ld=rb + ((si + 1) % 3))
                ldc = m + (si % 2) * 3
                C0 = oa.rand_matrix(rng, t, m, n, ld=ldc)
                if betas[(si + oi) % 3] == 0.0:
                    C0[:m] = np.nan  # beta == 0 must not read C (blas/level3_impl.h:64)
                c = C0.copy(order="F")
                al, be = alphas[(si + oi) % 3], betas[(si + oi) % 3]
                r = oa.call_gemm(getattr(L, t + "gemm_"), t, ta, tb, m, n, k, al, A, A.shape[0], B, B.shape[0], be, c, ldc)
                assert r == 0, eigen_b200.last_error()
                assert c[m:].tobytes() == C0[m:].tobytes(), "padding rows of C were touched"
                Cin = C0 if be != 0.0 else np.zeros_like(C0)
                fro, ratio = _check(t, ta, tb, m, n, k, al, be, A, B, Cin, c, ldc)
                worst = max(worst, ratio)
    L.b200blas_set_variant(0)
    print("worst gauge ratio", t, worst)


@pytest.mark.parametrize("t", list("sdcz"))
def test_quick_returns_and_k0(L, t):
    rng = np.random.default_rng(2)
    A = oa.rand_matrix(rng, t, 8, 8)
    C0 = oa.rand_matrix(rng, t, 8, 8)
    c = C0.copy(order="F")
    # k == 0: only the beta scaling (blas/level3_impl.h:62-69), alpha = NaN must not leak
    oa.call_gemm(getattr(L, t + "gemm_"), t, "N", "N", 8, 8, 0, np.nan, A, 8, A, 8, 1.3, c, 8)
    assert np.allclose(c, C0 * oa.NP_DTYPE[t](1.3), rtol=4 * oa.EPS[t])
    c = C0.copy(order="F")
    oa.call_gemm(getattr(L, t + "gemm_"), t, "N", "N", 8, 8, 0, 1.0, A, 8, A, 8, 0.0, c, 8)
    assert np.all(c == 0)
    # zero-sized products leave everything alone (test/product_extra.cpp:123-147)
    c = C0.copy(order="F")
    oa.call_gemm(getattr(L, t + "gemm_"), t, "N", "N", 0, 8, 8, 1.0, A, 8, A, 8, 0.0, c, 8)
    assert c.tobytes() == C0.tobytes()


@pytest.mark.parametrize("t", list("sdcz"))
def test_ones_known_answer(L, t):
    """test/product_extra.cpp:313-354 (Ones*Ones == k) through every op pair."""
    dt = oa.NP_DTYPE[t]
    for kk in (4, 300):
        a = np.ones((kk, kk), dtype=dt, order="F")
        for ta in "NTC":
            for tb in "NTC":
                c = np.zeros((kk, kk), dtype=dt, order="F")
                oa.call_gemm(getattr(L, t + "gemm_"), t, ta, tb, kk, kk, kk, 1.0, a, kk, a, kk, 0.0, c, kk)
                assert np.all(c == kk)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(HERE, "golden", "*.npz"))))
def test_committed_reference_outputs(L, path):
    """Outputs of the reference library itself (tests/golden, made by make_golden.py): ours must agree with them
    to the gauge ratio both sides are entitled to (2 x 16 eps G is the loosest the netlib criterion allows; we
    require 8)."""
    z = np.load(path)
    t = str(z["t"])
    m, n, k = [int(v) for v in z["mnk"]]
    ta, tb = str(z["ta"]), str(z["tb"])
    A, B, C0, Cref = [np.asfortranarray(z[x]) for x in ("A", "B", "C0", "Cref")]
    al, be = z["alpha"].item(), z["beta"].item()
    c = C0.copy(order="F")
    assert oa.call_gemm(getattr(L, t + "gemm_"), t, ta, tb, m, n, k, al, A, A.shape[0], B, B.shape[0], be, c, c.shape[0]) == 0
    assert c[m:].tobytes() == Cref[m:].tobytes()
    _, g = oa.hp_gemm(t, ta, tb, m, n, k, al, A, A.shape[0], B, B.shape[0], be, C0, C0.shape[0])
    assert (np.abs(c[:m] - Cref[:m]) / (oa.EPS[t] * g)).max() < 8.0


def test_device_pointer_api_and_streams(L):
    """Section 2 of include/b200blas.h: device-resident operands (torch tensors), asynchronous on a stream."""
    import torch
    torch.manual_seed(0)
    for t, dt in (("d", torch.float64), ("s", torch.float32), ("z", torch.complex128), ("c", torch.complex64)):
        m, n, k = 300, 260, 190
        A = (torch.rand(k, m, dtype=dt, device="cuda") * 2 - 1)  # column-major m x k == row-major k x m
        B = (torch.rand(n, k, dtype=dt, device="cuda") * 2 - 1)
        Cd = torch.ones(n, m, dtype=dt, device="cuda")
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            r = eigen_b200.gemm_dev(t, "N", "N", m, n, k, 1.0, A, m, B, k, 1.0, Cd, m)
        s.synchronize()
        assert r == 0
        An = np.asfortranarray(A.cpu().numpy().T)
        Bn = np.asfortranarray(B.cpu().numpy().T)
        C0 = np.ones((m, n), dtype=oa.NP_DTYPE[t], order="F")
        c = np.asfortranarray(Cd.cpu().numpy().T)
        _check(t, "N", "N", m, n, k, 1.0, 1.0, An, Bn, C0, c, m)


def test_lu_trailing_update_shape(L):
    """BASELINE config 5 (scaled down): A22 -= A21*A12 on sub-blocks of ONE matrix, lda=ldb=ldc (LU/PartialPivLU.h:492)."""
    rng = np.random.default_rng(7)
    N, bs = 1100, 64
    M = oa.rand_matrix(rng, "d", N, N)
    M0 = M.copy(order="F")
    A21, A12, A22 = M[bs:, :bs], M[:bs, bs:], M[bs:, bs:]
    mm = N - bs
    r = eigen_b200.gemm_host("d", "N", "N", mm, mm, bs, -1.0, A21, N, A12, N, 1.0, A22, N)
    assert r == 0
    assert np.array_equal(M[:bs], M0[:bs]) and np.array_equal(M[:, :bs], M0[:, :bs])
    ref, g = oa.hp_gemm("d", "N", "N", mm, mm, bs, -1.0, M0[bs:, :bs], N, M0[:bs, bs:], N, 1.0, M0[bs:, bs:], N)
    assert (np.abs(M[bs:, bs:] - ref) / (oa.EPS["d"] * g)).max() < 16.0


def test_concurrent_callers_are_safe(L):
    """The reference's ?gemm_ is re-entrant and Eigen may call it from many user threads: four threads
    issue host-pointer products of different types and shapes at once (ctypes releases the GIL); every result must
    equal the single-threaded one bit for bit."""
    import threading
    rng = np.random.default_rng(99)
    jobs = []
    for i, (t, m, n, k) in enumerate([("d", 700, 650, 900), ("s", 1100, 900, 1300), ("z", 300, 420, 510), ("c", 640, 512, 768),
                                      ("d", 129, 1031, 77), ("s", 333, 222, 111)]):
        A = oa.rand_matrix(rng, t, m, k)
        B = oa.rand_matrix(rng, t, k, n)
        C0 = oa.rand_matrix(rng, t, m, n)
        want = C0.copy(order="F")
        assert oa.call_gemm(getattr(L, t + "gemm_"), t, "N", "N", m, n, k, 1.0, A, m, B, k, 1.0, want, m) == 0
        jobs.append((t, m, n, k, A, B, C0, want))
    errors = []

    def worker(idx):
        for rep in range(3):
            t, m, n, k, A, B, C0, want = jobs[(idx + rep) % len(jobs)]
            c = C0.copy(order="F")
            r = oa.call_gemm(getattr(L, t + "gemm_"), t, "N", "N", m, n, k, 1.0, A, m, B, k, 1.0, c, m)
            if r != 0 or c.tobytes() != want.tobytes():
                errors.append((idx, rep, t, r))

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(4)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors


def test_pageable_large_operands_go_through_the_staging_ring(L):
    """Ordinary malloc'ed (pageable) matrices above the ring threshold: same result as the oracle, padding intact."""
    rng = np.random.default_rng(5)
    m, n, k = 1500, 1100, 1300          # every operand > 4 MiB -> pinned ring + copy workers + downloader thread
    for t in "ds":
        A = oa.rand_matrix(rng, t, m, k, ld=m + 5)
        B = oa.rand_matrix(rng, t, k, n, ld=k + 3)
        C0 = oa.rand_matrix(rng, t, m, n, ld=m + 7)
        c = C0.copy(order="F")
        assert oa.call_gemm(getattr(L, t + "gemm_"), t, "N", "T" if False else "N", m, n, k, 0.7, A, m + 5, B, k + 3, 1.3, c, m + 7) == 0
        assert c[m:].tobytes() == C0[m:].tobytes()
        _check(t, "N", "N", m, n, k, 0.7, 1.3, A, B, C0, c, m + 7, port_too=False)


# ---- round 2: the default LARGE-shape kernels on ragged shapes (VERDICT r1 "what's weak" 1 and 2) ------------------------
PAIR_SHAPES = [(1537, 2049, 1000), (2051, 1283, 4099), (777, 2600, 333)]


def _sample_rows(m, count=20):
    """First / last rows, both sides of every 128- and 256-row tile edge near the ends, and a spread in between."""
    edge = [0, 1, 127, 128, 255, 256, m - 257, m - 256, m - 129, m - 128, m - 2, m - 1]
    mid = list(np.linspace(0, m - 1, count).astype(int))
    return np.array(sorted({r for r in edge + mid if 0 <= r < m}), dtype=np.int32)


def _dev_matrix(torch, host):
    """An (ld x cols) F-ordered numpy array as a device tensor with the same column-major bytes."""
    return torch.from_numpy(np.ascontiguousarray(host.T)).cuda()


@pytest.mark.parametrize("t", list("sc"))
def test_cta_pair_kernels_ragged_shapes_all_ops(L, t):
    """The CTA-pair tcgen05 kernels (the default for m, n >= 512) on shapes that are multiples of nothing: ragged
    m / n / k (TMA out-of-bounds fill, row < m, ncol clipping, the 128-row B halves of a ragged last pair tile), all
    nine op pairs, ld = dim + 1, the xBLAT alpha / beta grid incl. beta = 0 with NaN in C, untouched padding rows --
    through ?gemm_ on host arrays AND b200blas_gemm_dev on device pointers (dblat3.f:395-675 semantics)."""
    import torch
    rng = np.random.default_rng(4321)
    cplx = t == "c"
    alphas = [1.0, (0.7 - 0.9j) if cplx else 0.7, -1.0]
    betas = [1.0, (1.3 - 1.1j) if cplx else 1.3, 0.0]
    ops = [(x, y) for x in "NTC" for y in "NTC"]
    worst = 0.0
    for si, (m, n, k) in enumerate(PAIR_SHAPES):
        rows = _sample_rows(m)
        for oi, (ta, tb) in enumerate(ops):
            ra, ca = (m, k) if ta == "N" else (k, m)
            rb, cb = (k, n) if tb == "N" else (n, k)
            A = oa.rand_matrix(rng, t, ra, ca, ld=ra + 1)
            B = oa.rand_matrix(rng, t, rb, cb, ld=rb + 1)
            ldc = m + 1
            C0 = oa.rand_matrix(rng, t, m, n, ld=ldc)
            al, be = alphas[(si + oi) % 3], betas[(si + 2 * oi) % 3]
            if be == 0.0:
                C0[:m] = np.nan  # beta == 0 must not read C (blas/level3_impl.h:64)
            Cin = C0 if be != 0.0 else np.zeros_like(C0)
            # (1) F77 entry, host arrays
            c = C0.copy(order="F")
            # (the host path cuts C into column slabs, so it reaches the pair kernels only for its widest slabs)
            assert oa.call_gemm(getattr(L, t + "gemm_"), t, ta, tb, m, n, k, al, A, ra + 1, B, rb + 1, be, c, ldc) == 0, eigen_b200.last_error()
            assert c[m:].tobytes() == C0[m:].tobytes(), "padding row of C was touched (host path)"
            _, ratio = _check(t, ta, tb, m, n, k, al, be, A, B, Cin, c, ldc, rows=rows, port_too=False)
            worst = max(worst, ratio)
            # (2) device pointers with the caller's odd leading dimensions
            dA, dB, dC = _dev_matrix(torch, A), _dev_matrix(torch, B), _dev_matrix(torch, C0)
            assert eigen_b200.gemm_dev(t, ta, tb, m, n, k, al, dA, ra + 1, dB, rb + 1, be, dC, ldc) == 0, eigen_b200.last_error()
            torch.cuda.synchronize()
            assert "pair" in eigen_b200.last_variant(), eigen_b200.last_variant()
            c2 = np.asfortranarray(dC.cpu().numpy().T)
            assert c2[m:].tobytes() == C0[m:].tobytes(), "padding row of C was touched (device path)"
            _, ratio = _check(t, ta, tb, m, n, k, al, be, A, B, Cin, c2, ldc, rows=rows, port_too=False)
            worst = max(worst, ratio)
    print("pair kernels, worst gauge ratio", t, worst)


@pytest.mark.parametrize("t", list("dz"))
def test_dmma_loader_modes_on_odd_device_pointers(L, t):
    """Device-pointer products whose operands are sub-blocks at ODD row / column offsets of matrices with ODD leading
    dimensions (what &lu(k+bs, k+bs) is): 8-byte aligned, not 16 -> the 8-byte cp.async loader (LD_DIM8); transposed
    operands -> LD_K; aligned even ones -> LD_DIM16.  More than 100 big tiles so the 128x64 configuration runs; every
    (A mode) x (B mode) combination of the real kernel, and the complex ones."""
    import torch
    rng = np.random.default_rng(2468)
    cplx = t == "z"
    # real: 17 x 21 = 357 big tiles = one full wave of 296 on the 128x64 tile + 61 tiles on the 64x32 tail launch (both
    # configurations see the odd pointers); complex: 13 x 19 = 247 big tiles, one partial wave of the big tile only
    m, n, k = (2100, 1300, 301) if not cplx else (801, 600, 203)
    want_variant = "dmma_z_64x32x16_w16x16_2cta_mbar" if cplx else "dmma_d_128x64x16_w32x32_2cta_mbar+tail"
    es = 16 if cplx else 8
    rows = _sample_rows(m, 12)
    # (row offset, column offset, ld parity): odd offsets + odd ld -> 8-byte aligned only; (0, 0, even) -> 16-byte aligned
    layouts = [(1, 3, 1), (0, 0, 0), (3, 1, 1)]
    al, be = ((0.7 - 0.9j), (1.3 - 1.1j)) if cplx else (0.7, 1.3)
    worst = 0.0
    for ta in "NTC":
        for tb in "NTC":
            for li, (ro, co, odd) in enumerate(layouts):
                if cplx and li == 2:
                    continue
                ra, ca = (m, k) if ta == "N" else (k, m)
                rb, cb = (k, n) if tb == "N" else (n, k)

                def parent(r, c_):
                    ld = r + ro + 2
                    ld += (ld % 2 == 0) if odd else (ld % 2 == 1)   # odd / even leading dimension
                    return oa.rand_matrix(rng, t, ld, c_ + co), ld

                PA, lda = parent(ra, ca)
                PB, ldb = parent(rb, cb)
                PC, ldc = parent(m, n)
                PC0 = PC.copy(order="F")
                dA, dB, dC = _dev_matrix(torch, PA), _dev_matrix(torch, PB), _dev_matrix(torch, PC)
                off = lambda ld: (ro + co * ld) * es  # noqa: E731
                r = eigen_b200.gemm_dev(t, ta, tb, m, n, k, al, dA.data_ptr() + off(lda), lda, dB.data_ptr() + off(ldb), ldb,
                                        be, dC.data_ptr() + off(ldc), ldc)
                assert r == 0, eigen_b200.last_error()
                torch.cuda.synchronize()
                assert eigen_b200.last_variant() == want_variant, eigen_b200.last_variant()
                got = np.asfortranarray(dC.cpu().numpy().T)
                # everything outside the m x n window is bit-identical
                mask = np.ones(PC0.shape, dtype=bool)
                mask[ro:ro + m, co:co + n] = False
                assert np.array_equal(got[mask], PC0[mask]), "elements outside the C window were modified"
                Asub = np.asfortranarray(PA[ro:, co:co + ca])
                Bsub = np.asfortranarray(PB[ro:, co:co + cb])
                Csub = np.asfortranarray(PC0[ro:, co:co + n])
                gsub = np.asfortranarray(got[ro:, co:co + n])
                _, ratio = _check(t, ta, tb, m, n, k, al, be, Asub, Bsub, Csub, gsub, Csub.shape[0], rows=rows, port_too=False)
                worst = max(worst, ratio)
    print("dmma loader modes, worst gauge ratio", t, worst)


@pytest.mark.parametrize("t", list("sc"))
def test_tf32x3_huge_and_infinite_inputs(L, t):
    """a finite value that rounds up to Inf under cvt.rna.tf32 must not overflow in the hi/lo split -- the
    tensor variant returns the same finite numbers as the SIMT variant (= the reference's fp32 FMA path).  A true +-Inf
    input cannot be carried through a split product (Inf * lo(b) is Inf * 0 = NaN whenever b is exactly representable);
    the documented behaviour (include/b200blas.h) is: the affected rows come back non-finite (Inf or NaN),
    every other row is unaffected."""
    rng = np.random.default_rng(77)
    m, n, k = 300, 280, 160
    fmax = float(np.finfo(np.float32).max)
    for case in ("huge", "inf"):
        A = oa.rand_matrix(rng, t, m, k)
        B = oa.rand_matrix(rng, t, k, n)
        if case == "huge":
            A[5, 7] = fmax * (1 - 2.0 ** -20)      # rounds UP to Inf under cvt.rna.tf32
            A[9, 3] = -fmax
            B[7, :] = 2.0 ** -40
            B[3, :] = 2.0 ** -40
        else:
            A[5, 7] = np.inf
            A[9, 3] = -np.inf
        C0 = np.zeros((m, n), dtype=oa.NP_DTYPE[t], order="F")
        outs = {}
        for variant in ("simt", "tf32x3"):
            L.b200blas_set_variant(eigen_b200.VARIANT[variant])
            c = C0.copy(order="F")
            assert oa.call_gemm(getattr(L, t + "gemm_"), t, "N", "N", m, n, k, 1.0, A, m, B, k, 0.0, c, m) == 0, eigen_b200.last_error()
            outs[variant] = c
        L.b200blas_set_variant(0)
        s, x = outs["simt"], outs["tf32x3"]
        if case == "huge":
            assert np.isfinite(s).all() and np.isfinite(x).all(), "a finite product overflowed in the split"
            assert np.abs(x - s).max() <= 64 * oa.EPS[t] * np.abs(s).max()
        else:
            bad = np.zeros(m, dtype=bool)
            bad[[5, 9]] = True
            assert not np.isfinite(x[bad]).any() and not np.isfinite(s[bad]).any()
            assert np.isfinite(x[~bad]).all()
            assert np.abs(x[~bad] - s[~bad]).max() <= 64 * oa.EPS[t] * np.abs(s[~bad]).max()


@pytest.mark.parametrize("head", ["staircase", "simple"])
def test_host_pipeline_head_with_chunked_a(L, head):
    """Host-operand products large enough for the multi-slab / chunked-A head of host.cu::run_host (k >= 4096, n >= 1024): the
    first slabs accumulate chunk by chunk while A is still on the bus, in whatever order the arrivals release the products
    (default: the staircase; B200BLAS_HOST_HEAD=simple: the four-slab head of the first schedule) -- N/T/C operands,
    alpha / beta != 1, beta = 0 with NaN in C, ld = dim + 1, ragged last chunk.  Same cases as tools/host_head_check.py, in a
    fresh process because the schedule is read once per process."""
    import subprocess
    import sys
    env = dict(os.environ)
    env.pop("B200BLAS_HOST_HEAD", None)
    if head == "simple":
        env["B200BLAS_HOST_HEAD"] = "simple"
    p = subprocess.run([sys.executable, os.path.join(HERE, "..", "tools", "host_head_check.py")], env=env, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, timeout=600, text=True)
    assert p.returncode == 0 and "all cases passed" in p.stdout, p.stdout[-3000:]
