/* include/b200blas.h -- C ABI of libb200blas.so, the sm_100a (NVIDIA B200) GEMM engine behind Eigen's BLAS seams.
 *
 * Section 1 is the drop-in boundary: exactly the symbols the reference binds for its dense matrix-matrix
 * product hot path.  Section 2 is the device-resident surface the benchmark uses (the reference has no device API for
 * this path).  Section 3 is the multi-GPU partitioner that sits behind sections 1 and 2.
 * Plain C types only -- no CUDA or torch types.
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
 *   exactly zero (the factorization is completed).  lapack/lu.cpp:14-42, Eigen/src/LU/PartialPivLU.h:361-496.
 *   No row limit: panels of up to 16384 rows run on the register-resident cluster kernel, taller ones on the cooperative
 *   slab kernel (shared memory, or a global-memory slab above ~120k rows of doubles); tests/test_gpu_lapack.py covers each.
 * Both run a right-looking blocked loop with one step of look-ahead on an internal high-priority stream; the call is still
 * ordered on the caller's stream (device pointers) / synchronous (host pointers).  Environment (forced variants for A/B
 * runs): B200BLAS_LOOKAHEAD=0, B200BLAS_POTRF=rec, B200BLAS_GETRF=rec, B200BLAS_GETF2=reg|slab|cluster, B200BLAS_TRSM=inv|subst. */
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

/* Non-finite inputs on the float tensor path (type S / C, variant tf32x3): every operand is split as x = hi + lo (3xTF32).  A
 * finite x always stays finite (values that would round up to Inf are truncated instead), NaN propagates, but a +-Inf input
 * makes the affected rows / columns of the result non-finite (Inf or NaN: Inf * lo(b) is Inf * 0 whenever b is exactly
 * representable in tf32) where the reference's fp32 FMA path returns +-Inf.  B200BLAS_VARIANT=simt keeps IEEE behaviour. */
int b200blas_gemm_dev(int type, char transa, char transb, int m, int n, int k, const void* alpha, const void* dA,
                      int64_t lda, const void* dB, int64_t ldb, const void* beta, void* dC, int64_t ldc,
                      void* stream, int variant);

/* Tensor contraction seam (replaces the LaunchKernels step of
 * unsupported/Eigen/CXX11/src/Tensor/TensorContractionCuda.h:1320-1390): out (m x n, column-major, leading dimension ldo)
 * = lhs * rhs, lhs(i, kk) = lhs[i * lhs_row_stride + kk * lhs_col_stride], rhs(kk, j) = rhs[kk * rhs_row_stride + j *
 * rhs_col_stride] (strides in elements, device pointers).  Returns 0 = done on `stream`, 1 = the views are not BLAS matrices
 * (no unit stride): the caller keeps its own path, -1 = error.  include/b200blas_eigen_tensor.h is the C++ glue the evaluator
 * calls. */
int b200blas_contract_dev(int type, int64_t m, int64_t n, int64_t k, const void* lhs, int64_t lhs_row_stride, int64_t lhs_col_stride,
                          const void* rhs, int64_t rhs_row_stride, int64_t rhs_col_stride, void* out, int64_t ldo, void* stream);

/* Tracing: B200BLAS_LOG=1 prints one line per product call on stderr (shape, operand residency, devices,
 * kernel variant, wall ms, bytes moved); B200BLAS_NVTX=1 opens one NVTX range per entry point (header-only nvtx3: nothing is
 * linked, nothing happens unless a profiler is attached). */

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

/* ------------------------------------------------------------------------------------------------------------
 * 3. Multi-GPU partitioner behind the same entry points (the B200 counterpart of parallelize_gemm,
 *    Eigen/src/Core/products/Parallelizer.h:85-157, invoked from GeneralMatrixMatrix.h:483-489: the caller writes
 *    one product and gets all workers).  ONE process drives N devices: C is cut into a pr x pc grid of tiles (2 -> 1x2,
 *    4 -> 2x2, 8 -> 2x4; k is never split, so there is no reduction), device (i,j) needs the row panel A_i and the column
 *    panel B_j.  Every k-chunk of a panel crosses the slow link ONCE (PCIe for host operands, the root's NVLink egress
 *    for device operands) to one "owner" device of its grid row / column and is relayed to the other devices of that
 *    row / column by peer-to-peer copy-engine transfers (no SMs involved); the products on the chunks that have landed
 *    overlap the transfers of the later ones; finished column sub-slabs of the C tiles flow back while the last
 *    products run (beta*C is folded in by a small kernel, or not read at all when beta == 0).
 *    Enabled by the environment variable B200BLAS_NGPUS=N or b200blas_set_devices(N); then ?gemm_ on host pointers,
 *    ?gemm_ on device pointers (operands resident on the current device = "root") and b200blas_gemm_dev use N devices
 *    for products of at least B200BLAS_MULTI_MIN_FLOPS (default 4.6e11 = 2*6144^3) flops; smaller ones stay on one GPU.
 *    B200BLAS_GRID=PRxPC / b200blas_set_grid override the grid.
 * ------------------------------------------------------------------------------------------------------------ */
int b200blas_set_devices(int n);              /* n <= 1: single device.  Returns the device count in effect (clamped to the visible devices) */
int b200blas_get_devices(void);
int b200blas_set_grid(int pr, int pc);        /* 0,0 restores the default grid; returns 0, or -1 if pr*pc does not match the device count */
/* page-lock / unlock caller memory so that every device can DMA it directly (cudaHostRegister, portable).  Optional:
 * pageable operands work (they are staged through pinned rings by copy threads), registered ones are faster. */
int b200blas_host_register(void* p, uint64_t bytes);
int b200blas_host_unregister(void* p);

/* The partition plan as data (pure host arithmetic; used by the executor, by tests/test_multi_plan.py -- which
 * interprets it on the CPU -- and by bench.py for reporting).  Regions are in elements of the scalar type, in the
 * column-major coordinates of their buffer. */
typedef struct {
  int loc;                 /* -1: the caller's operand ("origin": host memory, or the root device's memory); d >= 0: device d of the plan */
  int buf;                 /* origin: 0 = A, 1 = B, 2 = C.  device: 0 = panel A_i, 1 = panel B_j, 2 = product tile P, 3 = uploaded C tile,
                              4 + s = (root only) tile received from device s */
  int64_t r0, c0, rows, cols;
} b200blas_region;
enum { B200BLAS_STEP_COPY = 0, B200BLAS_STEP_GEMM = 1, B200BLAS_STEP_AXPBY = 2 };
typedef struct {
  int kind;                /* COPY: z := x.  GEMM: z := alpha*op(x)*op(y) + beta*z.  AXPBY: z := beta*z + alpha*x */
  int dev, stream;         /* issuing device (plan index) and stream slot: 0 fetch, 1 relay, 2 compute, 3 return, 4 fold */
  b200blas_region x, y, z;
  int opa, opb;            /* 0 = N, 1 = T, 2 = C */
  double alpha[2], beta[2];
  int nwait, wait[4];      /* steps (earlier in the list, other streams) that must have completed */
  int record;              /* some later step waits on this one */
} b200blas_step;
enum { B200BLAS_PLAN_MAXDEV = 8, B200BLAS_PLAN_MAXCHUNK = 64 };
typedef struct {
  int ndev, pr, pc, nchunks, ngroups, host_origin;
  int64_t row_cut[B200BLAS_PLAN_MAXDEV + 1], col_cut[B200BLAS_PLAN_MAXDEV + 1], k_cut[B200BLAS_PLAN_MAXCHUNK + 1];
  int group_first_chunk[B200BLAS_PLAN_MAXCHUNK + 1];
  int64_t ld[B200BLAS_PLAN_MAXDEV][4];      /* leading dimensions of the device buffers 0..3 (4 + s uses ld[s][2]) */
  int64_t elems[B200BLAS_PLAN_MAXDEV][4];   /* their sizes in elements (0: not used on that device) */
  int nsteps;
} b200blas_plan_info;
/* Builds the plan of one product.  host_origin: 1 = operands in host memory, 0 = resident on device 0 of the plan.
 * pr = pc = 0: default grid.  Returns the number of steps (written to steps[0 .. min(cap, nsteps)) ), or -1. */
int b200blas_multi_plan(int type, char transa, char transb, int64_t m, int64_t n, int64_t k, const double* alpha2,
                        const double* beta2, int64_t lda, int64_t ldb, int64_t ldc, int ndev, int pr, int pc,
                        int host_origin, b200blas_plan_info* info, b200blas_step* steps, int cap);

/* Pipe-peak micro-benchmarks (register-resident loops, no memory traffic) used as roofline denominators:
 * pipe 0 = FP64 DMMA (mma.sync.m8n8k4.f64), 1 = FP64 DFMA, 2 = FP32 FFMA, 3 = TF32 tcgen05.mma (dense).
 * Runs for about `millis` ms on the current device; returns TFLOP/s (<0 on failure). */
double b200blas_pipe_peak(int pipe, int millis);

#ifdef __cplusplus
}
#endif
#endif /* B200BLAS_H */
