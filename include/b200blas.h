/* include/b200blas.h -- C ABI of libb200blas.so, the sm_100a (NVIDIA B200) GEMM engine behind Eigen's BLAS seams.
 *
 * Section 1 is the drop-in boundary: exactly the symbols the reference binds for its dense matrix-matrix
 * product hot path.  Section 2 is the device-resident surface the benchmark and the multi-GPU driver use (the
 * reference has no device API for this path).  Plain C types only -- no CUDA or torch types.
 *
 * Reference interfaces replaced (paths relative to the PX4/eigen tree):
 *   Eigen/src/misc/blas.h:347-352                  prototypes of sgemm_/dgemm_/cgemm_/zgemm_
 *   Eigen/src/misc/blas.h:19, blas/xerbla.cpp:15   xerbla_ (weak, overridable)
 *   blas/level3_impl.h:12-76                       EIGEN_BLAS_FUNC(gemm): checks, info codes, beta pre-pass
 *   Eigen/src/Core/products/GeneralMatrixMatrix_BLAS.h:49-116   the EIGEN_USE_BLAS call site (beta == 1)
 */
#ifndef B200BLAS_H
#define B200BLAS_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------------------------
 * 1. Fortran-77 BLAS ABI (drop-in).  Everything by pointer, column-major, 32-bit ints; complex = interleaved
 *    (re,im) pairs; synchronous: C is complete in caller memory on return.  Semantics follow
 *    blas/level3_impl.h:47-69 exactly:
 *      info 1 bad transa | 2 bad transb | 3 m<0 | 4 n<0 | 5 k<0 | 8 lda<max(1,rows(A)) | 10 ldb<max(1,rows(B))
 *      | 13 ldc<max(1,m)  ->  xerbla_("xGEMM ", &info, 6), C untouched;
 *      m==0||n==0 -> return; beta==0 -> C window overwritten without being read; k==0 -> only the beta scaling;
 *      only the m x n window of C is written (rows m..ldc-1 of every column stay bit-identical).
 *    a/b/c may be host pointers (pageable or pinned; staged through pinned buffers and CUDA streams) or device
 *    pointers of the current CUDA device.  There is NO CPU fallback: if no sm_100 device or kernel is available
 *    the call reports through xerbla_ with info = -1 (CUDA failure) and leaves C untouched.
 * ------------------------------------------------------------------------------------------------------------ */
int sgemm_(const char* transa, const char* transb, const int* m, const int* n, const int* k, const float* alpha,
           const float* a, const int* lda, const float* b, const int* ldb, const float* beta, float* c,
           const int* ldc); /* replaces blas/single.cpp via blas/level3_impl.h:12 */
int dgemm_(const char* transa, const char* transb, const int* m, const int* n, const int* k, const double* alpha,
           const double* a, const int* lda, const double* b, const int* ldb, const double* beta, double* c,
           const int* ldc); /* replaces blas/double.cpp via blas/level3_impl.h:12 */
int cgemm_(const char* transa, const char* transb, const int* m, const int* n, const int* k, const float* alpha,
           const float* a, const int* lda, const float* b, const int* ldb, const float* beta, float* c,
           const int* ldc); /* replaces blas/complex_single.cpp via blas/level3_impl.h:12 */
int zgemm_(const char* transa, const char* transb, const int* m, const int* n, const int* k, const double* alpha,
           const double* a, const int* lda, const double* b, const int* ldb, const double* beta, double* c,
           const int* ldc); /* replaces blas/complex_double.cpp via blas/level3_impl.h:12 */
/* Rank-k updates, same kernels with a triangular tile mask.  Only the `uplo` triangle of C is read or
 * written.  Semantics of blas/level3_impl.h:357-433 (syrk) and :564-627 (herk): info 1 bad uplo | 2 bad trans ('C' is
 * invalid for csyrk_/zsyrk_, 'T' for ?herk_) | 3 n<0 | 4 k<0 | 7 lda<max(1,rows(A)) | 10 ldc<max(1,n); ?herk_ takes REAL
 * alpha and beta and stores the imaginary part of the diagonal as exactly zero whenever it writes the diagonal. */
int ssyrk_(const char* uplo, const char* trans, const int* n, const int* k, const float* alpha, const float* a,
           const int* lda, const float* beta, float* c, const int* ldc);   /* blas/single.cpp via level3_impl.h:357 */
int dsyrk_(const char* uplo, const char* trans, const int* n, const int* k, const double* alpha, const double* a,
           const int* lda, const double* beta, double* c, const int* ldc); /* blas/double.cpp via level3_impl.h:357 */
int csyrk_(const char* uplo, const char* trans, const int* n, const int* k, const float* alpha, const float* a,
           const int* lda, const float* beta, float* c, const int* ldc);   /* blas/complex_single.cpp via level3_impl.h:357 */
int zsyrk_(const char* uplo, const char* trans, const int* n, const int* k, const double* alpha, const double* a,
           const int* lda, const double* beta, double* c, const int* ldc); /* blas/complex_double.cpp via level3_impl.h:357 */
int cherk_(const char* uplo, const char* trans, const int* n, const int* k, const float* alpha, const float* a,
           const int* lda, const float* beta, float* c, const int* ldc);   /* blas/complex_single.cpp via level3_impl.h:564 */
int zherk_(const char* uplo, const char* trans, const int* n, const int* k, const double* alpha, const double* a,
           const int* lda, const double* beta, double* c, const int* ldc); /* blas/complex_double.cpp via level3_impl.h:564 */
/* The remaining level-3 routines, composites of the same kernels -- see eigen_b200/csrc/tri.cu.
 * ?trsm_: B := alpha * inv(op(A)) * B (side L) or alpha * B * inv(op(A)) (side R), A triangular; blas/level3_impl.h:78-178.
 * ?trmm_: B := alpha * op(A) * B or alpha * B * op(A); blas/level3_impl.h:183-284 (returns 1 like the reference).
 *   info 1 bad side | 2 bad uplo | 3 bad transa | 4 bad diag | 5 m<0 | 6 n<0 | 9 lda<max(1, side L ? m : n) | 11 ldb<max(1,m). */
int strsm_(const char* side, const char* uplo, const char* transa, const char* diag, const int* m, const int* n,
           const float* alpha, const float* a, const int* lda, float* b, const int* ldb);
int dtrsm_(const char* side, const char* uplo, const char* transa, const char* diag, const int* m, const int* n,
           const double* alpha, const double* a, const int* lda, double* b, const int* ldb);
int ctrsm_(const char* side, const char* uplo, const char* transa, const char* diag, const int* m, const int* n,
           const float* alpha, const float* a, const int* lda, float* b, const int* ldb);
int ztrsm_(const char* side, const char* uplo, const char* transa, const char* diag, const int* m, const int* n,
           const double* alpha, const double* a, const int* lda, double* b, const int* ldb);
int strmm_(const char* side, const char* uplo, const char* transa, const char* diag, const int* m, const int* n,
           const float* alpha, const float* a, const int* lda, float* b, const int* ldb);
int dtrmm_(const char* side, const char* uplo, const char* transa, const char* diag, const int* m, const int* n,
           const double* alpha, const double* a, const int* lda, double* b, const int* ldb);
int ctrmm_(const char* side, const char* uplo, const char* transa, const char* diag, const int* m, const int* n,
           const float* alpha, const float* a, const int* lda, float* b, const int* ldb);
int ztrmm_(const char* side, const char* uplo, const char* transa, const char* diag, const int* m, const int* n,
           const double* alpha, const double* a, const int* lda, double* b, const int* ldb);
/* ?symm_ / ?hemm_: C := alpha * A * B + beta * C (side L) or alpha * B * A + beta * C (side R), A symmetric / Hermitian with
 *   only its `uplo` triangle referenced; blas/level3_impl.h:287-355, :505-562.
 *   info 1 bad side | 2 bad uplo | 3 m<0 | 4 n<0 | 7 lda<max(1, side L ? m : n) | 9 ldb<max(1,m) | 12 ldc<max(1,m).
 * ?syr2k_ / ?her2k_: C.tri := alpha * op(A) op(B)^T + alpha * op(B) op(A)^T + beta * C.tri (her2k: ^H, conj(alpha) on the
 *   second term, REAL beta, real diagonal); blas/level3_impl.h:437-503, :631-700.
 *   info 1 bad uplo | 2 bad trans | 3 n<0 | 4 k<0 | 7 lda | 9 ldb | 12 ldc. */
int ssymm_(const char* side, const char* uplo, const int* m, const int* n, const float* alpha, const float* a, const int* lda,
            const float* b, const int* ldb, const float* beta, float* c, const int* ldc);
int dsymm_(const char* side, const char* uplo, const int* m, const int* n, const double* alpha, const double* a, const int* lda,
            const double* b, const int* ldb, const double* beta, double* c, const int* ldc);
int csymm_(const char* side, const char* uplo, const int* m, const int* n, const float* alpha, const float* a, const int* lda,
            const float* b, const int* ldb, const float* beta, float* c, const int* ldc);
int zsymm_(const char* side, const char* uplo, const int* m, const int* n, const double* alpha, const double* a, const int* lda,
            const double* b, const int* ldb, const double* beta, double* c, const int* ldc);
int chemm_(const char* side, const char* uplo, const int* m, const int* n, const float* alpha, const float* a, const int* lda,
            const float* b, const int* ldb, const float* beta, float* c, const int* ldc);
int zhemm_(const char* side, const char* uplo, const int* m, const int* n, const double* alpha, const double* a, const int* lda,
            const double* b, const int* ldb, const double* beta, double* c, const int* ldc);
int ssyr2k_(const char* uplo, const char* trans, const int* n, const int* k, const float* alpha, const float* a, const int* lda,
            const float* b, const int* ldb, const float* beta, float* c, const int* ldc);
int dsyr2k_(const char* uplo, const char* trans, const int* n, const int* k, const double* alpha, const double* a, const int* lda,
            const double* b, const int* ldb, const double* beta, double* c, const int* ldc);
int csyr2k_(const char* uplo, const char* trans, const int* n, const int* k, const float* alpha, const float* a, const int* lda,
            const float* b, const int* ldb, const float* beta, float* c, const int* ldc);
int zsyr2k_(const char* uplo, const char* trans, const int* n, const int* k, const double* alpha, const double* a, const int* lda,
            const double* b, const int* ldb, const double* beta, double* c, const int* ldc);
int cher2k_(const char* uplo, const char* trans, const int* n, const int* k, const float* alpha, const float* a, const int* lda,
            const float* b, const int* ldb, const float* beta, float* c, const int* ldc);
int zher2k_(const char* uplo, const char* trans, const int* n, const int* k, const double* alpha, const double* a, const int* lda,
            const double* b, const int* ldb, const double* beta, double* c, const int* ldc);
/* Device-resident blocked factorizations -- eigen_b200/csrc/lapack.cu.  LAPACK F77 ABI of the reference's
 * lapack/ module: the matrix is uploaded once (or already lives in HBM: device pointers are accepted for `a`; ipiv and
 * info are always HOST pointers), factored in place, and brought back once.
 * ?potrf_: A = L L^H (uplo 'L') or U^H U (uplo 'U'); only the `uplo` triangle is referenced / overwritten;
 *   info -1 bad uplo | -2 n<0 | -4 lda<max(1,n) (-> xerbla_("xPOTRF", &(-info), 6)); info = k > 0: the leading minor of
 *   order k is not positive definite.  lapack/cholesky.cpp:14-38, Eigen/src/Cholesky/LLT.h:299-360.
 * ?getrf_: P A = L U with partial pivoting, ipiv 1-based; info -1 m<0 | -2 n<0 | -4 lda<max(1,m); info = k > 0: U(k,k) is
 *   exactly zero (the factorization is completed).  lapack/lu.cpp:14-42, Eigen/src/LU/PartialPivLU.h:361-496. */
int spotrf_(const char* uplo, const int* n, float* a, const int* lda, int* info);
int dpotrf_(const char* uplo, const int* n, double* a, const int* lda, int* info);
int cpotrf_(const char* uplo, const int* n, float* a, const int* lda, int* info);
int zpotrf_(const char* uplo, const int* n, double* a, const int* lda, int* info);
int sgetrf_(const int* m, const int* n, float* a, const int* lda, int* ipiv, int* info);
int dgetrf_(const int* m, const int* n, double* a, const int* lda, int* ipiv, int* info);
int cgetrf_(const int* m, const int* n, float* a, const int* lda, int* ipiv, int* info);
int zgetrf_(const int* m, const int* n, double* a, const int* lda, int* ipiv, int* info);
/* Weak default prints "Eigen BLAS ERROR #<info>: <name>" like blas/xerbla.cpp:15-19; applications and testers
 * override it by defining their own xerbla_. */
int xerbla_(const char* name, int* info, int len);

/* ------------------------------------------------------------------------------------------------------------
 * 2. Device-resident API (inputs already in HBM).  type: 0 = float, 1 = double, 2 = complex float,
 *    3 = complex double.  alpha/beta are HOST pointers to one scalar of that type.  stream is a cudaStream_t
 *    passed as void* (NULL = legacy default stream); the call is asynchronous on that stream.
 *    Same argument checks / info codes / quick returns as section 1.  Returns 0, or the xerbla_ return value.
 * ------------------------------------------------------------------------------------------------------------ */
enum { B200BLAS_S = 0, B200BLAS_D = 1, B200BLAS_C = 2, B200BLAS_Z = 3 };
/* kernel variants (b200blas_gemm_dev's variant argument, env B200BLAS_VARIANT=auto|simt|dmma|tf32x3) */
enum { B200BLAS_AUTO = 0, B200BLAS_SIMT = 1, B200BLAS_DMMA = 2, B200BLAS_TF32X3 = 3 };

int b200blas_gemm_dev(int type, char transa, char transb, int m, int n, int k, const void* alpha, const void* dA,
                      int64_t lda, const void* dB, int64_t ldb, const void* beta, void* dC, int64_t ldc,
                      void* stream, int variant);

/* introspection used by tests and bench.py */
int b200blas_version(void);
int b200blas_device_ok(void);                 /* 1 if the current device is sm_100 and the kernels loaded */
const char* b200blas_last_error(void);        /* last CUDA error string seen by this thread, "" if none */
const char* b200blas_last_variant(void);      /* variant name of this thread's last product, e.g. "dmma_128x128x16" */
uint64_t b200blas_kernel_launches(void);      /* kernels launched by this library since load (all threads) */
void b200blas_set_variant(int variant);       /* process-wide override; B200BLAS_AUTO restores the heuristic */
/* host staging statistics of the last F77 call on this thread (bytes moved over PCIe) */
void b200blas_last_transfer(uint64_t* h2d_bytes, uint64_t* d2h_bytes);
/* release cached device workspaces / pinned staging buffers */
void b200blas_release(void);

/* Pipe-peak micro-benchmarks (register-resident loops, no memory traffic) used as roofline denominators:
 * pipe 0 = FP64 DMMA (mma.sync.m8n8k4.f64), 1 = FP64 DFMA, 2 = FP32 FFMA, 3 = TF32 tcgen05.mma (dense).
 * Runs for about `millis` ms on the current device; returns TFLOP/s (<0 on failure). */
double b200blas_pipe_peak(int pipe, int millis);

#ifdef __cplusplus
}
#endif
#endif /* B200BLAS_H */
