"""Pins the CPU oracle (oracle/liboracle.so) before anything trusts it -- CPU only.

Sources of truth, in order:
  1. the reference itself, compiled here from /root/reference into oracle/_ref (skipped where _ref is absent);
  2. committed outputs of that reference in tests/golden/*.npz (made by tests/golden/make_golden.py);
  3. the known answers / properties the reference's own tests hold for this path:
     Ones*Ones == k (test/product_extra.cpp:313-354), xBLAT3 ratio < 16 and argument preservation
     (blas/testing/dblat3.f:395-675,2508-2627), error exits (dblat3.f:1889-1972).
"""
import ctypes as C
import glob
import os

import numpy as np
import pytest

import oracle_api as oa

P = oa.port()
needs_ref = pytest.mark.skipif(not oa.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
HERE = os.path.dirname(os.path.abspath(__file__))


def _chk1(t, fn):
    rep = oa.Blat3Report()
    P.oracle_blat3_chk1(oa.TYPES[t], C.cast(fn, C.c_void_p), C.byref(rep))
    return rep


@pytest.mark.parametrize("t", list("sdcz"))
def test_port_passes_xblat3(t):
    rep = _chk1(t, getattr(P, "oracle_%sgemm_" % t))
    assert rep.ncalls == 17496  # 6*6*6 dims x 9 op pairs x 9 (alpha,beta)
    assert not rep.fatal, rep.msg
    assert rep.errmax < 16.0
    log = C.create_string_buffer(4096)
    assert P.oracle_blat3_chke(oa.TYPES[t], C.cast(getattr(P, "oracle_%sgemm_" % t), C.c_void_p),
                               C.cast(P.oracle_set_xerbla, C.c_void_p), log, 4096) == 0, log.value


@needs_ref
@pytest.mark.parametrize("t", list("sdcz"))
def test_reference_passes_the_restated_xblat3(t):
    """The checker itself is validated on the reference's own library: same call count, passes, and the port's
    ERRMAX equals the reference's (both round identically on these shapes)."""
    rep = _chk1(t, getattr(oa.ref_blas(), t + "gemm_"))
    assert rep.ncalls == 17496 and not rep.fatal and rep.errmax < 16.0, rep.msg
    rep_p = _chk1(t, getattr(P, "oracle_%sgemm_" % t))
    assert abs(rep.errmax - rep_p.errmax) < 1.0
    log = C.create_string_buffer(4096)
    assert P.oracle_blat3_chke(oa.TYPES[t], C.cast(getattr(oa.ref_blas(), t + "gemm_"), C.c_void_p), None, log, 4096) == 0


def test_blat3_generator_first_values():
    # DBEG: i = 7; i = i*891 mod 1000 -> 237 -> (237-500)/1001
    P.oracle_blat3_reset()
    assert abs(P.oracle_blat3_dbeg() - (237 - 500) / 1001.0) < 1e-15
    assert abs(P.oracle_blat3_dbeg() - ((237 * 891) % 1000 - 500) / 1001.0) < 1e-15


@pytest.mark.parametrize("t", list("sdcz"))
def test_ones_times_ones_known_answer(t):
    """test/product_extra.cpp:313-354: Ones(4,4)*Ones(4,4) == 4 for every storage-order mix."""
    dt = oa.NP_DTYPE[t]
    for ta in "NTC":
        for tb in "NTC":
            a = np.ones((4, 4), dtype=dt, order="F")
            c = np.zeros((4, 4), dtype=dt, order="F")
            oa.call_gemm(getattr(P, "oracle_%sgemm_" % t), t, ta, tb, 4, 4, 4, 1.0, a, 4, a, 4, 0.0, c, 4)
            assert np.all(c == 4)


@needs_ref
def test_blocking_sizes_match_reference():
    l1, l2, l3 = oa.sync_cache_sizes()
    S = oa.ref_shim()
    for t in "sdcz":
        for threads in (1, 8):
            for (m, n, k) in ((2048, 2048, 2048), (16384, 16384, 16384), (16384, 16384, 256), (8192, 8192, 8192),
                              (4096, 4096, 4096), (37, 29, 53), (300, 7, 5000), (1, 1, 1), (47, 47, 47), (48, 1, 1)):
                a = [C.c_long(k), C.c_long(m), C.c_long(n)]
                b = [C.c_long(k), C.c_long(m), C.c_long(n)]
                S.ref_blocking_sizes(oa.TYPES[t], *[C.byref(x) for x in a], threads)
                P.oracle_blocking_sizes(oa.TYPES[t], *[C.byref(x) for x in b], threads)
                assert [x.value for x in a] == [x.value for x in b], (t, threads, m, n, k)
        mr, nr, lp = C.c_int(), C.c_int(), C.c_int()
        mr2, nr2, lp2 = C.c_int(), C.c_int(), C.c_int()
        S.ref_gebp_traits(oa.TYPES[t], C.byref(mr), C.byref(nr), C.byref(lp))
        P.oracle_gebp_traits(oa.TYPES[t], C.byref(mr2), C.byref(nr2), C.byref(lp2))
        assert (mr.value, nr.value, lp.value) == (mr2.value, nr2.value, lp2.value)


@needs_ref
@pytest.mark.parametrize("t", list("sdz"))
def test_packed_panel_layouts_are_byte_identical(t):
    """gemm_pack_lhs / gemm_pack_rhs (GeneralBlockPanelKernel.h:1688-2105) vs the port, byte for byte."""
    S = oa.ref_shim()
    rng = np.random.default_rng(5)
    dt = oa.NP_DTYPE[t]
    for (rows, depth) in ((1, 1), (3, 5), (12, 8), (13, 7), (29, 11), (64, 16), (77, 33)):
        for order in (0, 1):
            src = oa.rand_matrix(rng, t, rows if order == 0 else depth, depth if order == 0 else rows)
            stride = src.shape[0]
            for side in ("lhs", "rhs"):
                out_r = _aligned_zeros(rows * depth, dt)  # the reference packs with aligned SIMD stores
                out_p = _aligned_zeros(rows * depth, dt)
                # for the rhs the roles are (depth x cols): reuse "rows" as cols with the transposed order
                o = order if side == "lhs" else 1 - order
                args = (oa._ptr(src), stride, depth, rows, o)
                if t == "z":
                    getattr(S, "ref_pack_%s_z" % side)(oa._ptr(out_r), *args, 0)
                else:
                    getattr(S, "ref_pack_%s_%s" % (side, t))(oa._ptr(out_r), *args)
                getattr(P, "oracle_pack_%s" % side)(oa.TYPES[t], oa._ptr(out_p), *args, 0)
                assert out_r.tobytes() == out_p.tobytes(), (t, rows, depth, order, side)


def _aligned_zeros(count, dt, align=64):
    raw = np.zeros(count * np.dtype(dt).itemsize + align, dtype=np.uint8)
    off = (-raw.ctypes.data) % align
    return raw[off:off + count * np.dtype(dt).itemsize].view(dt)


def _packet(t):
    return {"s": 8, "d": 4, "c": 4, "z": 2}[t]


@needs_ref
@pytest.mark.parametrize("t", list("sdcz"))
def test_port_vs_reference_blas_bitwise(t):
    """Against the reference's own ?gemm_ (oracle/_ref/libeigen_blas_ref.so, blas/level3_impl.h:12-76) with the
    host's cache sizes handed to the port: bit-for-bit equal on every element computed by the packet paths of
    gebp_kernel (rows below the last full SIMD packet); the few scalar-tail rows agree to a 2-ulp gauge ratio."""
    oa.sync_cache_sizes()
    rng = np.random.default_rng(11)
    RB = oa.ref_blas()
    cplx = t in "cz"
    al = (0.7 - 0.9j) if cplx else 0.7
    be = (1.3 - 1.1j) if cplx else 1.3
    shapes = [(64, 64, 64), (37, 29, 53), (200, 180, 700), (515, 130, 401), (96, 7, 1000), (8, 300, 9)]
    for (m, n, k) in shapes:
        for ta in "NTC":
            for tb in "NTC":
                A = oa.rand_matrix(rng, t, m if ta == "N" else k, k if ta == "N" else m, ld=(m if ta == "N" else k) + 3)
                B = oa.rand_matrix(rng, t, k if tb == "N" else n, n if tb == "N" else k, ld=(k if tb == "N" else n) + 1)
                C0 = oa.rand_matrix(rng, t, m, n, ld=m + 2)
                c1, c2 = C0.copy(order="F"), C0.copy(order="F")
                oa.call_gemm(getattr(RB, t + "gemm_"), t, ta, tb, m, n, k, al, A, A.shape[0], B, B.shape[0], be, c1, m + 2)
                oa.call_gemm(getattr(P, "oracle_%sgemm_" % t), t, ta, tb, m, n, k, al, A, A.shape[0], B, B.shape[0], be, c2, m + 2)
                assert c1[m:].tobytes() == c2[m:].tobytes() == C0[m:].tobytes()  # ld padding untouched by both
                body = (m // _packet(t)) * _packet(t)
                assert c1[:body].tobytes() == c2[:body].tobytes(), (t, m, n, k, ta, tb)
                if body < m:
                    ref, g = oa.hp_gemm(t, ta, tb, m, n, k, al, A, A.shape[0], B, B.shape[0], be, C0, m + 2)
                    ratio = np.abs(c1[body:m] - c2[body:m]) / (oa.EPS[t] * g[body:m])
                    assert ratio.max() < 2.0


@needs_ref
@pytest.mark.parametrize("t", list("sdcz"))
def test_port_omp_vs_reference_openmp_gebp(t):
    """parallelize_gemm + the OpenMP branch (Parallelizer.h:85-157, GeneralMatrixMatrix.h:83-152) through Eigen's
    public expression API vs the port's threaded driver: same kc => same per-element rounding on packet rows."""
    oa.sync_cache_sizes()
    rng = np.random.default_rng(3)
    S = oa.ref_shim()
    m, n, k = 256, 384, 900
    A = oa.rand_matrix(rng, t, m, k)
    B = oa.rand_matrix(rng, t, k, n)
    C0 = oa.rand_matrix(rng, t, m, n)
    c1, c2 = C0.copy(order="F"), C0.copy(order="F")
    dt = oa.NP_DTYPE[t]
    al, be = np.array([1.0], dtype=dt), np.array([1.0], dtype=dt)
    getattr(S, "ref_eigen_gemm_" + t)(b"N", b"N", m, n, k, oa._ptr(al), oa._ptr(A), m, oa._ptr(B), k, oa._ptr(be), oa._ptr(c1), m, 4)
    P.oracle_gemm_omp(oa.TYPES[t], b"N", b"N", m, n, k, oa._ptr(al), oa._ptr(A), m, oa._ptr(B), k, oa._ptr(be), oa._ptr(c2), m, 4)
    ref, g = oa.hp_gemm(t, "N", "N", m, n, k, 1.0, A, m, B, k, 1.0, C0, m)
    # alpha = 1 is folded differently by the expression API (scaleAndAddTo), so allow a 2-ulp gauge ratio overall
    assert (np.abs(c1 - c2) / (oa.EPS[t] * g)).max() < 2.0
    assert (np.abs(c2 - ref) / (oa.EPS[t] * g)).max() < 16.0


def test_parallel_partition_restates_parallelize_gemm():
    """Parallelizer.h:108-155: threads = min(nbThreads, cols/nr, m*n*k/50000); column slabs multiples of 4, row
    slices multiples of mr; the last thread takes the remainder."""
    arr = lambda: (C.c_long * 64)()  # noqa: E731
    c0, nc, r0, nr = arr(), arr(), arr(), arr()
    T = P.oracle_parallel_partition(oa.TYPES["d"], 16384, 16384, 16384, 8, 0, c0, nc, r0, nr)
    assert T == 8
    assert list(nc[:8]) == [2048] * 8 and list(c0[:8]) == [2048 * i for i in range(8)]
    assert list(nr[:7]) == [2040] * 7 and nr[7] == 16384 - 7 * 2040  # 2048 rounded down to mr = 12
    T = P.oracle_parallel_partition(oa.TYPES["d"], 10, 7, 10, 8, 0, c0, nc, r0, nr)
    assert T == 1  # 700 flop-units < 50000
    T = P.oracle_parallel_partition(oa.TYPES["d"], 1000, 10, 1000, 8, 0, c0, nc, r0, nr)
    assert T == 2 and sum(nc[:2]) == 10


def test_hp_reference_is_exact_on_integers():
    rng = np.random.default_rng(0)
    for t in "sdcz":
        dt = oa.NP_DTYPE[t]
        A = rng.integers(-8, 8, size=(9, 13)).astype(dt)
        B = rng.integers(-8, 8, size=(13, 5)).astype(dt)
        if t in "cz":
            A = A + 1j * rng.integers(-8, 8, size=(9, 13)).astype(dt)
            B = B + 1j * rng.integers(-8, 8, size=(13, 5)).astype(dt)
        A, B = np.asfortranarray(A), np.asfortranarray(B)
        Cm = np.zeros((9, 5), dtype=dt, order="F")
        ref, g = oa.hp_gemm(t, "N", "N", 9, 5, 13, 1.0, A, 9, B, 13, 0.0, Cm, 9)
        assert np.array_equal(ref, (A.astype(np.complex128) @ B.astype(np.complex128)) if t in "cz" else A.astype(np.float64) @ B.astype(np.float64))
        ref2, _ = oa.hp_gemm(t, "C", "T", 9, 5, 13, 1.0, np.asfortranarray(A.conj().T), 13, np.asfortranarray(B.T), 5, 0.0, Cm, 9, rows=[8, 0])
        assert np.array_equal(ref2, ref[[8, 0]])


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(HERE, "golden", "*.npz"))))
def test_port_vs_committed_reference_outputs(path):
    """tests/golden/*.npz hold inputs and the outputs of the reference library computed in the build container
    (tests/golden/make_golden.py).  The port must reproduce them: bitwise on packet rows when the cache sizes
    recorded in the fixture are used, within a 2-ulp gauge ratio elsewhere."""
    z = np.load(path)
    t = str(z["t"])
    P.oracle_set_cache_sizes(*[int(v) for v in z["cache"]])
    m, n, k = [int(v) for v in z["mnk"]]
    ta, tb = str(z["ta"]), str(z["tb"])
    A, B, C0, Cref = [np.asfortranarray(z[x]) for x in ("A", "B", "C0", "Cref")]
    al, be = z["alpha"].item(), z["beta"].item()
    c = C0.copy(order="F")
    oa.call_gemm(getattr(P, "oracle_%sgemm_" % t), t, ta, tb, m, n, k, al, A, A.shape[0], B, B.shape[0], be, c, c.shape[0])
    body = (m // _packet(t)) * _packet(t)
    assert c[:body].tobytes() == Cref[:body].tobytes()
    assert c[m:].tobytes() == Cref[m:].tobytes()
    ref, g = oa.hp_gemm(t, ta, tb, m, n, k, al, A, A.shape[0], B, B.shape[0], be, C0, C0.shape[0])
    assert (np.abs(c[:m] - Cref[:m]) / (oa.EPS[t] * g)).max() < 2.0
