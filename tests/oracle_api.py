"""ctypes bindings for the CPU checkers under oracle/ (TEST INFRASTRUCTURE ONLY).

* ``port``  -> oracle/liboracle.so          plain-C restatement (oracle/gebp_port.c, hp_ref.c, blat3_port.c)
* ``ref``   -> oracle/_ref/*.so             the unmodified reference compiled by oracle/Makefile (may be absent)

The product package ``eigen_b200`` never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

TYPES = {"s": 0, "d": 1, "c": 2, "z": 3}
NP_DTYPE = {"s": np.float32, "d": np.float64, "c": np.complex64, "z": np.complex128}
EPS = {"s": 2.0 ** -23, "d": 2.0 ** -52, "c": 2.0 ** -23, "z": 2.0 ** -52}

_i = C.c_int
_ip = C.POINTER(C.c_int)
_vp = C.c_void_p
_cp = C.c_char_p

GEMM_ARGTYPES = [_cp, _cp, _ip, _ip, _ip, _vp, _vp, _ip, _vp, _ip, _vp, _vp, _ip]
RANKK_ARGTYPES = [_cp, _cp, _ip, _ip, _vp, _vp, _ip, _vp, _vp, _ip]
RANKK_NAMES = ["ssyrk_", "dsyrk_", "csyrk_", "zsyrk_", "cherk_", "zherk_"]
TRI_ARGTYPES = [_cp, _cp, _cp, _cp, _ip, _ip, _vp, _vp, _ip, _vp, _ip]
TRI_NAMES = [t + r for r in ("trsm_", "trmm_") for t in "sdcz"]
ABC_ARGTYPES = [_cp, _cp, _ip, _ip, _vp, _vp, _ip, _vp, _ip, _vp, _vp, _ip]
SYMM_NAMES = ["ssymm_", "dsymm_", "csymm_", "zsymm_", "chemm_", "zhemm_"]
R2K_NAMES = ["ssyr2k_", "dsyr2k_", "csyr2k_", "zsyr2k_", "cher2k_", "zher2k_"]


class Blat3Report(C.Structure):
    _fields_ = [("ncalls", C.c_int), ("errmax", C.c_double), ("fatal", C.c_int), ("bad_param", C.c_int),
                ("msg", C.c_char * 200)]


def _build_port():
    so = os.path.join(ORACLE_DIR, "liboracle.so")
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("gebp_port.c", "gebp_impl.h", "hp_ref.c", "blat3_port.c", "rankk_port.c", "level3_port.c", "lapack_port.c", "oracle.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "port"], stdout=subprocess.DEVNULL)
    return so


POTRF_ARGTYPES = [_cp, _ip, _vp, _ip, _ip]
GETRF_ARGTYPES = [_ip, _ip, _vp, _ip, _ip, _ip]


def _bind_lapack(lib, prefix):
    for t in "sdcz":
        f = getattr(lib, prefix + t + "potrf_")
        f.argtypes = POTRF_ARGTYPES
        f.restype = _i
        f = getattr(lib, prefix + t + "getrf_")
        f.argtypes = GETRF_ARGTYPES
        f.restype = _i


_ref_lapack = None


def have_ref_lapack():
    return os.path.exists(os.path.join(ORACLE_DIR, "_ref", "libeigen_lapack_ref.so"))


def ref_lapack():
    """The reference's own ?potrf_/?getrf_ (lapack/cholesky.cpp, lapack/lu.cpp via oracle/ref_lapack_shim.cpp)."""
    global _ref_lapack
    if _ref_lapack is None:
        port()
        lib = C.CDLL(os.path.join(ORACLE_DIR, "_ref", "libeigen_lapack_ref.so"), mode=C.RTLD_LOCAL)
        _bind_lapack(lib, "")
        _ref_lapack = lib
    return _ref_lapack


def call_potrf(fn, uplo, n, a, lda):
    """F77 ?potrf_; returns info."""
    info = C.c_int(12345)
    ints = [C.c_int(n), C.c_int(lda)]
    fn(uplo.encode(), C.byref(ints[0]), _ptr(a), C.byref(ints[1]), C.byref(info))
    return info.value


def call_getrf(fn, m, n, a, lda):
    """F77 ?getrf_; returns (ipiv (1-based, length min(m,n)), info)."""
    info = C.c_int(12345)
    ipiv = np.zeros(max(min(m, n), 1), dtype=np.int32)
    ints = [C.c_int(m), C.c_int(n), C.c_int(lda)]
    fn(C.byref(ints[0]), C.byref(ints[1]), _ptr(a), C.byref(ints[2]), ipiv.ctypes.data_as(_ip), C.byref(info))
    return ipiv[:min(m, n)], info.value


def _bind_level3(lib, prefix):
    for nm in TRI_NAMES:
        f = getattr(lib, prefix + nm)
        f.argtypes = TRI_ARGTYPES
        f.restype = _i
    for nm in SYMM_NAMES + R2K_NAMES:
        f = getattr(lib, prefix + nm)
        f.argtypes = ABC_ARGTYPES
        f.restype = _i


_port = None
_ref_blas = None
_ref_shim = None


def port():
    """liboracle.so, loaded RTLD_GLOBAL so that its xerbla_ (the xBLAT3 tester's) interposes the weak ones."""
    global _port
    if _port is None:
        lib = C.CDLL(_build_port(), mode=C.RTLD_GLOBAL)
        for sfx in "sdcz":
            f = getattr(lib, "oracle_%sgemm_" % sfx)
            f.argtypes = GEMM_ARGTYPES
            f.restype = _i
        for nm in RANKK_NAMES:
            f = getattr(lib, "oracle_" + nm)
            f.argtypes = RANKK_ARGTYPES
            f.restype = _i
        _bind_level3(lib, "oracle_")
        _bind_lapack(lib, "oracle_")
        lib.oracle_xerbla_expect.argtypes = [_cp, _i]
        lib.oracle_xerbla_result.restype = _i
        lib.oracle_gemm_omp.argtypes = [_i, C.c_char, C.c_char, _i, _i, _i, _vp, _vp, _i, _vp, _i, _vp, _vp, _i, _i]
        lib.oracle_hp_gemm.argtypes = [_i, C.c_char, C.c_char, _i, _i, _i, _vp, _vp, _i, _vp, _i, _vp, _vp, _i, _vp, _i, _vp, _vp]
        lib.oracle_hp_gemm.restype = None
        lib.oracle_blat3_chk1.argtypes = [_i, _vp, C.POINTER(Blat3Report)]
        lib.oracle_blat3_chk1.restype = None
        lib.oracle_blat3_chke.argtypes = [_i, _vp, _vp, C.c_char_p, _i]
        lib.oracle_blat3_chke.restype = _i
        lib.oracle_set_cache_sizes.argtypes = [C.c_long] * 3
        lib.oracle_blocking_sizes.argtypes = [_i, C.POINTER(C.c_long), C.POINTER(C.c_long), C.POINTER(C.c_long), _i]
        lib.oracle_gebp_traits.argtypes = [_i, _ip, _ip, _ip]
        lib.oracle_pack_lhs.argtypes = [_i, _vp, _vp, C.c_long, C.c_long, C.c_long, _i, _i]
        lib.oracle_pack_rhs.argtypes = [_i, _vp, _vp, C.c_long, C.c_long, C.c_long, _i, _i]
        lib.oracle_parallel_partition.argtypes = [_i, C.c_long, C.c_long, C.c_long, _i, _i] + [C.POINTER(C.c_long)] * 4
        lib.oracle_parallel_partition.restype = _i
        lib.oracle_blat3_dbeg.restype = C.c_double
        _port = lib
    return _port


def have_ref():
    d = os.path.join(ORACLE_DIR, "_ref")
    return os.path.exists(os.path.join(d, "libeigen_blas_ref.so")) and os.path.exists(os.path.join(d, "libeigen_gebp_omp.so"))


def ref_blas():
    """The reference's own blas/ library (sgemm_/dgemm_/cgemm_/zgemm_ ... of blas/level3_impl.h)."""
    global _ref_blas
    if _ref_blas is None:
        port()  # tester xerbla_ first
        lib = C.CDLL(os.path.join(ORACLE_DIR, "_ref", "libeigen_blas_ref.so"), mode=C.RTLD_LOCAL)
        for sfx in "sdcz":
            f = getattr(lib, "%sgemm_" % sfx)
            f.argtypes = GEMM_ARGTYPES
            f.restype = _i
        for nm in RANKK_NAMES:
            f = getattr(lib, nm)
            f.argtypes = RANKK_ARGTYPES
            f.restype = _i
        _bind_level3(lib, "")
        _ref_blas = lib
    return _ref_blas


def ref_shim():
    """Eigen's expression API + OpenMP gebp path (oracle/ref_eigen_shim.cpp compiled against the reference)."""
    global _ref_shim
    if _ref_shim is None:
        lib = C.CDLL(os.path.join(ORACLE_DIR, "_ref", "libeigen_gebp_omp.so"), mode=C.RTLD_LOCAL)
        for sfx in "sdcz":
            f = getattr(lib, "ref_eigen_gemm_%s" % sfx)
            f.argtypes = [C.c_char, C.c_char, _i, _i, _i, _vp, _vp, _i, _vp, _i, _vp, _vp, _i, _i]
            f.restype = None
        lib.ref_cache_sizes.argtypes = [C.POINTER(C.c_long)] * 3
        lib.ref_blocking_sizes.argtypes = [_i, C.POINTER(C.c_long), C.POINTER(C.c_long), C.POINTER(C.c_long), _i]
        lib.ref_gebp_traits.argtypes = [_i, _ip, _ip, _ip]
        for n in ("ref_pack_lhs_d", "ref_pack_rhs_d", "ref_pack_lhs_s", "ref_pack_rhs_s"):
            getattr(lib, n).argtypes = [_vp, _vp, C.c_long, C.c_long, C.c_long, _i]
        for n in ("ref_pack_lhs_z", "ref_pack_rhs_z"):
            getattr(lib, n).argtypes = [_vp, _vp, C.c_long, C.c_long, C.c_long, _i, _i]
        _ref_shim = lib
    return _ref_shim


def sync_cache_sizes():
    """Give the port the cache sizes the reference detected on this host, so kc/mc/nc (and rounding) coincide."""
    l1, l2, l3 = C.c_long(), C.c_long(), C.c_long()
    ref_shim().ref_cache_sizes(C.byref(l1), C.byref(l2), C.byref(l3))
    port().oracle_set_cache_sizes(l1.value, l2.value, l3.value)
    return l1.value, l2.value, l3.value


def _ptr(a):
    return a.ctypes.data_as(_vp) if a is not None else None


def call_gemm(fn, t, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc):
    """Call an F77-ABI gemm: everything by pointer.  a/b/c are 1-D or F-ordered numpy arrays of the right dtype."""
    dt = NP_DTYPE[t]
    al = np.array([alpha], dtype=dt)
    be = np.array([beta], dtype=dt)
    ints = [C.c_int(v) for v in (m, n, k, lda, ldb, ldc)]
    return fn(ta.encode(), tb.encode(), C.byref(ints[0]), C.byref(ints[1]), C.byref(ints[2]), _ptr(al), _ptr(a),
              C.byref(ints[3]), _ptr(b), C.byref(ints[4]), _ptr(be), _ptr(c), C.byref(ints[5]))


def hp_gemm(t, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, rows=None):
    """Long-double reference.  Returns (C_ref, G) for the selected rows; C_ref is float64/complex128."""
    dt = NP_DTYPE[t]
    al = np.array([alpha], dtype=dt)
    be = np.array([beta], dtype=dt)
    cplx = t in "cz"
    if rows is None:
        ridx, nr = None, m
    else:
        ridx = np.ascontiguousarray(rows, dtype=np.int32)
        nr = len(ridx)
    out = np.zeros((nr, n), dtype=np.complex128 if cplx else np.float64, order="F")
    g = np.zeros((nr, n), dtype=np.float64, order="F")
    port().oracle_hp_gemm(TYPES[t], ta.encode(), tb.encode(), m, n, k, _ptr(al), _ptr(a), lda, _ptr(b), ldb, _ptr(be),
                          _ptr(c), ldc, _ptr(ridx), nr, _ptr(out), _ptr(g))
    return out, g


def rand_matrix(rng, t, rows, cols, ld=None):
    """Uniform [-1,1] like Eigen's setRandom (MathFunctions.h:628-637); returns an (ld x cols) F-ordered array."""
    ld = rows if ld is None else ld
    dt = NP_DTYPE[t]
    x = rng.uniform(-1, 1, size=(max(ld, 1), cols))
    if t in "cz":
        x = x + 1j * rng.uniform(-1, 1, size=(max(ld, 1), cols))
    return np.asfortranarray(x.astype(dt))


def call_rankk(fn, name, uplo, trans, n, k, alpha, a, lda, beta, c, ldc):
    """Call an F77-ABI ?syrk_/?herk_.  `name` like "dsyrk_" / "zherk_"; herk takes REAL alpha and beta."""
    t = name[0]
    herk = "herk" in name
    dt = NP_DTYPE[t]
    rdt = np.float32 if t in "sc" else np.float64
    al = np.array([alpha], dtype=rdt if herk else dt)
    be = np.array([beta], dtype=rdt if herk else dt)
    ints = [C.c_int(v) for v in (n, k, lda, ldc)]
    return fn(uplo.encode(), trans.encode(), C.byref(ints[0]), C.byref(ints[1]), _ptr(al), _ptr(a), C.byref(ints[2]),
              _ptr(be), _ptr(c), C.byref(ints[3]))


def hp_rankk(name, uplo, trans, n, k, alpha, a, lda, beta, c, ldc):
    """Long-double reference of a rank-k update (full n x n result and gauge; compare on the `uplo` triangle only)."""
    t = name[0]
    herk = "herk" in name
    other = ("C" if herk else "T")
    ta, tb = ("N", other) if trans in "Nn" else (other, "N")
    return hp_gemm(t, ta, tb, n, n, k, alpha, a, lda, a, lda, beta, c, ldc)


def tri_mask(n, uplo):
    i, j = np.indices((n, n))
    return (i <= j) if uplo in "Uu" else (i >= j)


def call_tri(fn, name, side, uplo, trans, diag, m, n, alpha, a, lda, b, ldb):
    """Call an F77-ABI ?trsm_/?trmm_."""
    dt = NP_DTYPE[name[0]]
    al = np.array([alpha], dtype=dt)
    ints = [C.c_int(v) for v in (m, n, lda, ldb)]
    return fn(side.encode(), uplo.encode(), trans.encode(), diag.encode(), C.byref(ints[0]), C.byref(ints[1]), _ptr(al), _ptr(a),
              C.byref(ints[2]), _ptr(b), C.byref(ints[3]))


def call_abc(fn, name, c1, c2, d1, d2, alpha, a, lda, b, ldb, beta, c, ldc):
    """Call an F77-ABI ?symm_/?hemm_ (c1 = side, c2 = uplo, d1 = m, d2 = n) or ?syr2k_/?her2k_ (uplo, trans, n, k);
    her2k takes a REAL beta."""
    t = name[0]
    dt = NP_DTYPE[t]
    rdt = np.float32 if t in "sc" else np.float64
    al = np.array([alpha], dtype=dt)
    be = np.array([beta], dtype=rdt if "her2k" in name else dt)
    ints = [C.c_int(v) for v in (d1, d2, lda, ldb, ldc)]
    return fn(c1.encode(), c2.encode(), C.byref(ints[0]), C.byref(ints[1]), _ptr(al), _ptr(a), C.byref(ints[2]), _ptr(b),
              C.byref(ints[3]), _ptr(be), _ptr(c), C.byref(ints[4]))


def dense_triangular(a, n, uplo, diag):
    """The n x n triangular matrix ?trsm_/?trmm_ see in `a` (other triangle zero, unit diagonal if diag == 'U')."""
    t = np.array(a[:n, :n], order="F")
    t = np.triu(t) if uplo in "Uu" else np.tril(t)
    if diag in "Uu":
        np.fill_diagonal(t, 1)
    return np.asfortranarray(t)


def dense_symmetric(a, n, uplo, herm):
    """The n x n symmetric / Hermitian matrix ?symm_/?hemm_ see in the `uplo` triangle of `a`."""
    t = np.array(a[:n, :n], order="F")
    strict = np.triu(t, 1) if uplo in "Uu" else np.tril(t, -1)
    d = np.diagonal(t).copy()
    if herm:
        d = d.real.astype(t.dtype)
    full = strict + (strict.conj().T if herm else strict.T) + np.diag(d)
    return np.asfortranarray(full.astype(t.dtype))


def make_triangular(rng, t, n, ld, scale=None):
    """A well-conditioned random triangular operand: off-diagonal uniform[-1,1] * scale, |diagonal| in [1, 2]
    (the netlib generator also shifts the diagonal, dblat3.f DMAKE); both triangles are filled."""
    scale = (2.0 / max(n, 2)) if scale is None else scale
    a = rand_matrix(rng, t, n, n, ld=ld)
    a[:n] = a[:n] * np.asarray(scale, dtype=a.real.dtype)
    d = np.diagonal(a[:n]).copy()
    d = np.where(np.abs(d) > 0, d / np.maximum(np.abs(d), 1e-30), 1) * (1 + rng.uniform(0, 1, size=n))
    a[np.arange(n), np.arange(n)] = d.astype(a.dtype)
    return a
