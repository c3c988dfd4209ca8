"""eigen_b200 -- B200-native (sm_100a) GEMM engine behind Eigen's BLAS seams.

The product is the C-ABI shared library ``eigen_b200/libb200blas.so`` (sources in ``eigen_b200/csrc``, interface in
``include/b200blas.h``).  This module is the thin Python host mirror used by tests, ``bench.py`` and the multi-GPU
driver: it binds the same entry points the reference's callers bind (``sgemm_/dgemm_/cgemm_/zgemm_`` of
``blas/level3_impl.h:12-76``, called by ``Eigen/src/Core/products/GeneralMatrixMatrix_BLAS.h:103``) plus the
device-resident API.  There is no CPU fallback: if the library is missing or no sm_100 device is present the calls
raise.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200blas.so")
CSRC = os.path.join(_HERE, "csrc")

TYPE_CODE = {"s": 0, "d": 1, "c": 2, "z": 3}
VARIANT = {"auto": 0, "simt": 1, "dmma": 2, "tf32x3": 3}

# every symbol include/b200blas.h declares
EXPORTS = ["sgemm_", "dgemm_", "cgemm_", "zgemm_", "ssyrk_", "dsyrk_", "csyrk_", "zsyrk_", "cherk_", "zherk_",
           "strsm_", "dtrsm_", "ctrsm_", "ztrsm_", "strmm_", "dtrmm_", "ctrmm_", "ztrmm_", "ssymm_", "dsymm_", "csymm_", "zsymm_",
           "chemm_", "zhemm_", "ssyr2k_", "dsyr2k_", "csyr2k_", "zsyr2k_", "cher2k_", "zher2k_",
           "spotrf_", "dpotrf_", "cpotrf_", "zpotrf_", "sgetrf_", "dgetrf_", "cgetrf_", "zgetrf_", "xerbla_", "b200blas_gemm_dev", "b200blas_version",
           "b200blas_device_ok", "b200blas_last_error", "b200blas_last_variant", "b200blas_kernel_launches",
           "b200blas_set_variant", "b200blas_last_transfer", "b200blas_release", "b200blas_pipe_peak",
           "b200blas_set_devices", "b200blas_get_devices", "b200blas_set_grid", "b200blas_host_register",
           "b200blas_host_unregister", "b200blas_multi_plan", "b200blas_contract_dev"]


class Region(C.Structure):
    """b200blas_region (include/b200blas.h section 3)"""
    _fields_ = [("loc", C.c_int), ("buf", C.c_int), ("r0", C.c_int64), ("c0", C.c_int64), ("rows", C.c_int64), ("cols", C.c_int64)]


class Step(C.Structure):
    """b200blas_step"""
    _fields_ = [("kind", C.c_int), ("dev", C.c_int), ("stream", C.c_int), ("x", Region), ("y", Region), ("z", Region),
                ("opa", C.c_int), ("opb", C.c_int), ("alpha", C.c_double * 2), ("beta", C.c_double * 2),
                ("nwait", C.c_int), ("wait", C.c_int * 4), ("record", C.c_int)]


PLAN_MAXDEV, PLAN_MAXCHUNK = 8, 64


class PlanInfo(C.Structure):
    """b200blas_plan_info"""
    _fields_ = [("ndev", C.c_int), ("pr", C.c_int), ("pc", C.c_int), ("nchunks", C.c_int), ("ngroups", C.c_int), ("host_origin", C.c_int),
                ("row_cut", C.c_int64 * (PLAN_MAXDEV + 1)), ("col_cut", C.c_int64 * (PLAN_MAXDEV + 1)), ("k_cut", C.c_int64 * (PLAN_MAXCHUNK + 1)),
                ("group_first_chunk", C.c_int * (PLAN_MAXCHUNK + 1)), ("ld", (C.c_int64 * 4) * PLAN_MAXDEV), ("elems", (C.c_int64 * 4) * PLAN_MAXDEV),
                ("nsteps", C.c_int)]

_lib = None


def build(verbose=False):
    """Compile libb200blas.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", CSRC, "-j8"], stdout=out)
    return LIB_PATH


def _needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    for f in os.listdir(CSRC):
        if f.endswith((".cu", ".cuh", ".h")) and os.path.getmtime(os.path.join(CSRC, f)) > t:
            return True
    return os.path.getmtime(os.path.join(_HERE, "..", "include", "b200blas.h")) > t


def lib():
    """The loaded C-ABI library (built on first use if nvcc is available; raises otherwise)."""
    global _lib
    if _lib is not None:
        return _lib
    if _needs_build():
        if not os.path.exists("/usr/local/cuda/bin/nvcc") and not os.path.exists(LIB_PATH):
            raise RuntimeError("libb200blas.so is missing and nvcc is unavailable; there is no CPU fallback")
        if os.path.exists("/usr/local/cuda/bin/nvcc"):
            build()
    L = C.CDLL(LIB_PATH, mode=C.RTLD_LOCAL)
    ip, vp, cp, i = C.POINTER(C.c_int), C.c_void_p, C.c_char_p, C.c_int
    for name in ("sgemm_", "dgemm_", "cgemm_", "zgemm_"):
        f = getattr(L, name)
        f.argtypes = [cp, cp, ip, ip, ip, vp, vp, ip, vp, ip, vp, vp, ip]
        f.restype = i
    for name in ("ssyrk_", "dsyrk_", "csyrk_", "zsyrk_", "cherk_", "zherk_"):
        f = getattr(L, name)
        f.argtypes = [cp, cp, ip, ip, vp, vp, ip, vp, vp, ip]
        f.restype = i
    for name in EXPORTS:
        if name[1:] in ("trsm_", "trmm_"):
            f = getattr(L, name)
            f.argtypes = [cp, cp, cp, cp, ip, ip, vp, vp, ip, vp, ip]
            f.restype = i
        elif name[1:] in ("symm_", "hemm_", "syr2k_", "her2k_"):
            f = getattr(L, name)
            f.argtypes = [cp, cp, ip, ip, vp, vp, ip, vp, ip, vp, vp, ip]
            f.restype = i
    for t in "sdcz":
        f = getattr(L, t + "potrf_")
        f.argtypes = [cp, ip, vp, ip, ip]
        f.restype = i
        f = getattr(L, t + "getrf_")
        f.argtypes = [ip, ip, vp, ip, ip, ip]
        f.restype = i
    L.b200blas_gemm_dev.argtypes = [i, C.c_char, C.c_char, i, i, i, vp, vp, C.c_int64, vp, C.c_int64, vp, vp,
                                    C.c_int64, vp, i]
    L.b200blas_gemm_dev.restype = i
    L.b200blas_last_error.restype = cp
    L.b200blas_last_variant.restype = cp
    L.b200blas_kernel_launches.restype = C.c_uint64
    L.b200blas_set_variant.argtypes = [i]
    L.b200blas_last_transfer.argtypes = [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.b200blas_pipe_peak.argtypes = [i, i]
    L.b200blas_pipe_peak.restype = C.c_double
    L.b200blas_set_devices.argtypes = [i]
    L.b200blas_set_grid.argtypes = [i, i]
    L.b200blas_host_register.argtypes = [vp, C.c_uint64]
    L.b200blas_host_unregister.argtypes = [vp]
    L.b200blas_contract_dev.argtypes = [i, C.c_int64, C.c_int64, C.c_int64, vp, C.c_int64, C.c_int64, vp, C.c_int64, C.c_int64, vp, C.c_int64, vp]
    L.b200blas_multi_plan.argtypes = [i, C.c_char, C.c_char, C.c_int64, C.c_int64, C.c_int64, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                      C.c_int64, C.c_int64, C.c_int64, i, i, i, i, C.POINTER(PlanInfo), C.POINTER(Step), i]
    _lib = L
    return L


def require_device():
    L = lib()
    if not L.b200blas_device_ok():
        raise RuntimeError("eigen_b200 needs an sm_100 (B200) CUDA device; there is no CPU fallback")
    return L


def gemm_host(t, transa, transb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc):
    """F77-ABI call ``?gemm_`` on host numpy arrays (column-major storage); ``c`` is updated in place."""
    import numpy as np
    dt = {"s": np.float32, "d": np.float64, "c": np.complex64, "z": np.complex128}[t]
    L = require_device()
    al = np.array([alpha], dtype=dt)
    be = np.array([beta], dtype=dt)
    ints = [C.c_int(v) for v in (m, n, k, lda, ldb, ldc)]
    ptr = lambda x: x.ctypes.data_as(C.c_void_p)  # noqa: E731
    return getattr(L, t + "gemm_")(transa.encode(), transb.encode(), C.byref(ints[0]), C.byref(ints[1]),
                                    C.byref(ints[2]), ptr(al), ptr(a), C.byref(ints[3]), ptr(b), C.byref(ints[4]),
                                    ptr(be), ptr(c), C.byref(ints[5]))


def gemm_dev(t, transa, transb, m, n, k, alpha, dA, lda, dB, ldb, beta, dC, ldc, stream=None, variant="auto"):
    """Device-resident product on raw device pointers (ints) or torch tensors; asynchronous on ``stream``."""
    import numpy as np
    dt = {"s": np.float32, "d": np.float64, "c": np.complex64, "z": np.complex128}[t]
    L = require_device()
    al = np.array([alpha], dtype=dt)
    be = np.array([beta], dtype=dt)

    def dp(x):
        return C.c_void_p(x if isinstance(x, int) else x.data_ptr())

    if stream is None:
        try:
            import torch
            stream = torch.cuda.current_stream().cuda_stream
        except Exception:
            stream = 0
    r = L.b200blas_gemm_dev(TYPE_CODE[t], transa.encode(), transb.encode(), m, n, k, al.ctypes.data_as(C.c_void_p),
                            dp(dA), lda, dp(dB), ldb, be.ctypes.data_as(C.c_void_p), dp(dC), ldc,
                            C.c_void_p(stream), VARIANT[variant])
    return r


def last_variant():
    return lib().b200blas_last_variant().decode()


def last_error():
    return lib().b200blas_last_error().decode()


def kernel_launches():
    return int(lib().b200blas_kernel_launches())


def pipe_peak(pipe, millis=300):
    """TFLOP/s of a register-resident loop on pipe 0=DMMA fp64, 1=DFMA, 2=FFMA, 3=tcgen05 tf32."""
    return float(require_device().b200blas_pipe_peak(pipe, millis))


def set_devices(n):
    """Use n GPUs of this box for large products behind ?gemm_ / gemm_dev (include/b200blas.h section 3); returns the count in effect."""
    return int(lib().b200blas_set_devices(int(n)))


def multi_plan(t, transa, transb, m, n, k, alpha, beta, ndev, grid=(0, 0), host_origin=False, cap=4096):
    """The partition plan of one product as data: (PlanInfo, [Step, ...]).  Pure host arithmetic -- works without a GPU."""
    al = (C.c_double * 2)(complex(alpha).real, complex(alpha).imag)
    be = (C.c_double * 2)(complex(beta).real, complex(beta).imag)
    info = PlanInfo()
    steps = (Step * max(cap, 1))()
    ns = lib().b200blas_multi_plan(TYPE_CODE[t], transa.encode(), transb.encode(), m, n, k, al, be, 0, 0, 0, ndev, grid[0], grid[1],
                                   1 if host_origin else 0, C.byref(info), steps, cap)
    if ns < 0 or (cap > 0 and ns > cap):
        raise ValueError("b200blas_multi_plan failed (%d)" % ns)
    return info, [steps[i] for i in range(min(ns, cap))]
