"""The multi-GPU partitioner behind ?gemm_ / b200blas_gemm_dev (include/b200blas.h section 3) on hardware -- run with -m gpu.

On a box with ONE B200 (the driver's GPU test tier) the plan devices are mapped onto the same physical GPU
(B200BLAS_MULTI_VIRTUAL=1): every plan device still has its own streams, events and panel buffers, so the executor, its
dependencies and the fold / return path are exercised exactly as on 2, 4 or 8 GPUs (peer copies become device-to-device
copies).  With more GPUs visible (tools/gpu_multi_r02.sh, gpurun --gpus N) the same tests use real peers.
Checks: small-integer operands make every summation order exact, so the whole C must EQUAL numpy's result; uniform[-1,1]
operands are checked on sampled rows against the long-double oracle (blas/testing/dblat3.f:2587-2596 gauge ratio < 16).
Reference: the parallel split inside the product call, Eigen/src/Core/products/Parallelizer.h:85-157.
"""
import ctypes as C
import os

import numpy as np
import pytest

import eigen_b200
import oracle_api as oa

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def L():
    lib = eigen_b200.require_device()
    old = {k: os.environ.get(k) for k in ("B200BLAS_MULTI_VIRTUAL", "B200BLAS_MULTI_MIN_FLOPS")}
    os.environ["B200BLAS_MULTI_VIRTUAL"] = "1"
    os.environ["B200BLAS_MULTI_MIN_FLOPS"] = "0"
    yield lib
    lib.b200blas_set_devices(1)
    lib.b200blas_set_grid(0, 0)
    for k, v in old.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


def _ints(rng, t, rows, cols, ld):
    x = rng.integers(-2, 3, size=(ld, cols)).astype(np.float64)
    if t in "cz":
        x = x + 1j * rng.integers(-2, 3, size=(ld, cols))
    return np.asfortranarray(x.astype(oa.NP_DTYPE[t]))


def _op(x, o):
    return x if o == "N" else x.T if o == "T" else x.conj().T


OPS = [(x, y) for x in "NTC" for y in "NTC"]
SHAPES = [(700, 900, 1300), (1537, 2049, 1000), (520, 3000, 2600)]


@pytest.mark.parametrize("ndev,grid", [(2, (0, 0)), (4, (0, 0)), (8, (0, 0)), (8, (1, 8)), (4, (4, 1))])
@pytest.mark.parametrize("t", list("dscz"))
def test_host_operands_exact_on_integer_matrices(L, t, ndev, grid):
    """?gemm_ on host arrays (pageable numpy memory, ld > dim) through N plan devices: exact equality with numpy on the whole
    matrix, padding rows untouched, beta == 0 never reads C."""
    assert L.b200blas_set_devices(ndev) == ndev
    assert L.b200blas_set_grid(*grid) == 0
    rng = np.random.default_rng(100 * ndev + ord(t))
    cplx = t in "cz"
    for si, (m, n, k) in enumerate(SHAPES):
        ta, tb = OPS[(si * 4 + ndev + ord(t)) % 9]
        ra, ca = (m, k) if ta == "N" else (k, m)
        rb, cb = (k, n) if tb == "N" else (n, k)
        A, B, C0 = _ints(rng, t, ra, ca, ra + 1), _ints(rng, t, rb, cb, rb + 3), _ints(rng, t, m, n, m + 2)
        al = (2 - 1j) if cplx else 2.0
        be = [(-3 + 2j) if cplx else -3.0, 0.0, 1.0][(si + ndev) % 3]
        if be == 0.0:
            C0[:m] = np.nan
        c = C0.copy(order="F")
        l0 = eigen_b200.kernel_launches()
        assert oa.call_gemm(getattr(L, t + "gemm_"), t, ta, tb, m, n, k, al, A, ra + 1, B, rb + 3, be, c, m + 2) == 0, eigen_b200.last_error()
        assert eigen_b200.kernel_launches() - l0 >= ndev, "the product did not run on every plan device"
        want = al * (_op(A[:ra].astype(np.complex128 if cplx else np.float64), ta) @ _op(B[:rb].astype(np.complex128 if cplx else np.float64), tb))
        if be != 0.0:
            want = want + be * C0[:m]
        assert np.array_equal(c[:m], want.astype(oa.NP_DTYPE[t])), (t, ta, tb, m, n, k, ndev, grid)
        assert c[m:].tobytes() == C0[m:].tobytes(), "padding rows of C were touched"
        h2d, d2h = C.c_uint64(), C.c_uint64()
        L.b200blas_last_transfer(C.byref(h2d), C.byref(d2h))
        es = np.dtype(oa.NP_DTYPE[t]).itemsize
        assert h2d.value == es * (m * k + k * n + (m * n if be != 0.0 else 0)), "every operand byte crosses PCIe exactly once"
        assert d2h.value == es * m * n


@pytest.mark.parametrize("t", list("ds"))
def test_host_operands_pinned_and_large_pageable(L, t):
    """Registered (page-locked) caller memory is DMA'd directly; large pageable operands go through the pinned rings and the
    downloader thread.  Uniform[-1,1] operands against the long-double oracle on sampled rows."""
    assert L.b200blas_set_devices(4) == 4
    assert L.b200blas_set_grid(0, 0) == 0
    rng = np.random.default_rng(7)
    m, n, k = 2304, 2100, 2500      # tiles of 1152 x 1050 doubles = 9.7 MB: sub-slab returns above the 1 MiB ring threshold
    A = oa.rand_matrix(rng, t, m, k, ld=m + 5)
    B = oa.rand_matrix(rng, t, k, n, ld=k + 3)
    C0 = oa.rand_matrix(rng, t, m, n, ld=m + 7)
    rows = np.array([0, 1, 1151, 1152, 1153, m - 1], dtype=np.int32)
    ref, g = oa.hp_gemm(t, "N", "N", m, n, k, 0.7, A, m + 5, B, k + 3, 1.3, C0, m + 7, rows=rows)
    for pinned in (False, True):
        c = C0.copy(order="F")
        if pinned:
            for x in (A, B, c):
                assert L.b200blas_host_register(C.c_void_p(x.ctypes.data), x.nbytes) == 0
        try:
            assert oa.call_gemm(getattr(L, t + "gemm_"), t, "N", "N", m, n, k, 0.7, A, m + 5, B, k + 3, 1.3, c, m + 7) == 0, eigen_b200.last_error()
        finally:
            if pinned:
                for x in (A, B, c):
                    L.b200blas_host_unregister(C.c_void_p(x.ctypes.data))
        assert c[m:].tobytes() == C0[m:].tobytes()
        ratio = (np.abs(c[rows] - ref) / (oa.EPS[t] * g)).max()
        assert ratio < 16.0, (t, pinned, ratio)


@pytest.mark.parametrize("ndev,grid", [(2, (0, 0)), (4, (0, 0)), (8, (0, 0)), (8, (1, 8))])
def test_device_resident_root_operands(L, ndev, grid):
    """Operands resident on the current device (the root): b200blas_gemm_dev is asynchronous on the caller's stream; three
    calls are queued back to back without a host synchronisation (the panel buffers of call i+1 must wait for call i),
    C += A*B each time; plus dgemm_ on device pointers.  Exact on integer matrices."""
    assert L.b200blas_set_devices(ndev) == ndev
    assert L.b200blas_set_grid(*grid) == 0
    rng = np.random.default_rng(5 + ndev)
    for t, (m, n, k), (ta, tb) in (("d", (1537, 2049, 2100), ("N", "N")), ("z", (700, 900, 1300), ("C", "T")), ("s", (2048, 1024, 4096), ("T", "N"))):
        ra, ca = (m, k) if ta == "N" else (k, m)
        rb, cb = (k, n) if tb == "N" else (n, k)
        A, B, C0 = _ints(rng, t, ra, ca, ra + 1), _ints(rng, t, rb, cb, rb + 1), _ints(rng, t, m, n, m + 1)
        dA, dB, dC = (torch.from_numpy(np.ascontiguousarray(x.T)).cuda() for x in (A, B, C0))
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            for _ in range(3):
                assert eigen_b200.gemm_dev(t, ta, tb, m, n, k, 1.0, dA, ra + 1, dB, rb + 1, 1.0, dC, m + 1) == 0, eigen_b200.last_error()
        s.synchronize()
        prod = _op(A[:ra].astype(np.complex128), ta) @ _op(B[:rb].astype(np.complex128), tb)
        want = (C0[:m] + 3 * prod).astype(oa.NP_DTYPE[t])
        got = np.asfortranarray(dC.cpu().numpy().T)
        assert np.array_equal(got[:m], want), (t, ndev, grid)
        assert got[m:].tobytes() == C0[m:].tobytes()
        # F77 entry on device pointers, beta = 0 over NaN
        dC.fill_(float("nan"))
        fn = getattr(L, t + "gemm_")
        ints = [C.c_int(v) for v in (m, n, k, ra + 1, rb + 1, m + 1)]
        al = np.array([2.0], dtype=oa.NP_DTYPE[t])
        be = np.array([0.0], dtype=oa.NP_DTYPE[t])
        assert fn(ta.encode(), tb.encode(), C.byref(ints[0]), C.byref(ints[1]), C.byref(ints[2]), oa._ptr(al), C.c_void_p(dA.data_ptr()), C.byref(ints[3]),
                  C.c_void_p(dB.data_ptr()), C.byref(ints[4]), oa._ptr(be), C.c_void_p(dC.data_ptr()), C.byref(ints[5])) == 0
        got = np.asfortranarray(dC.cpu().numpy().T)
        assert np.array_equal(got[:m], (2 * prod).astype(oa.NP_DTYPE[t]))


def test_small_products_stay_on_one_device(L):
    """Below B200BLAS_MULTI_MIN_FLOPS the call takes the single-device path (no panel traffic for a product that fits one GPU's
    launch latency)."""
    assert L.b200blas_set_devices(4) == 4
    os.environ["B200BLAS_MULTI_MIN_FLOPS"] = "4.6e11"
    try:
        rng = np.random.default_rng(3)
        A, B, c = oa.rand_matrix(rng, "d", 300, 300), oa.rand_matrix(rng, "d", 300, 300), oa.rand_matrix(rng, "d", 300, 300)
        l0 = eigen_b200.kernel_launches()
        assert oa.call_gemm(L.dgemm_, "d", "N", "N", 300, 300, 300, 1.0, A, 300, B, 300, 0.0, c, 300) == 0
        assert eigen_b200.kernel_launches() - l0 == 1
    finally:
        os.environ["B200BLAS_MULTI_MIN_FLOPS"] = "0"
