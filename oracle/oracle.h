/* oracle/oracle.h -- CPU ORACLE: TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the reference's dense GEMM hot path (PX4/eigen, Eigen 3.3.90), used as the checker
 * for the sm_100a library in eigen_b200/.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this; the product library never links, imports or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_pin.py checks this port against the reference itself
 * (oracle/_ref/libeigen_blas_ref.so and libeigen_gebp_omp.so, compiled from /root/reference by oracle/Makefile)
 * on the xBLAT3 grid, on Eigen's own known-answer (Ones*Ones == k, test/product_extra.cpp:313-354), and against
 * committed outputs of the reference in tests/golden/ (made by tests/golden/make_golden.py).
 *
 * Every function cites the reference file:line it follows (paths relative to /root/reference).
 */
#ifndef ORACLE_H
#define ORACLE_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

/* scalar type codes used across the oracle API */
enum { ORACLE_S = 0, ORACLE_D = 1, ORACLE_C = 2, ORACLE_Z = 3 };
/* op codes, blas/common.h:24-42 */
enum { ORACLE_NOTR = 0, ORACLE_TR = 1, ORACLE_ADJ = 2, ORACLE_INVALID = 0xff };

/* ---- cache model + blocking (GeneralBlockPanelKernel.h:39-78, 92-308) ------------------------------------ */
void oracle_set_cache_sizes(long l1, long l2, long l3);
void oracle_get_cache_sizes(long* l1, long* l2, long* l3);
/* gebp_traits<S,S>::{mr,nr,LhsProgress} for an AVX2+FMA build (GeneralBlockPanelKernel.h:369-380,618-619) */
void oracle_gebp_traits(int type, int* mr, int* nr, int* lhs_progress);
/* computeProductBlockingSizes: k,m,n in = problem, out = kc,mc,nc */
void oracle_blocking_sizes(int type, long* k, long* m, long* n, int num_threads);

/* ---- packing (GeneralBlockPanelKernel.h:1688-2105).  order 0 = ColMajor source, 1 = RowMajor source ------- */
void oracle_pack_lhs(int type, void* blockA, const void* lhs, long stride, long depth, long rows, int order, int conj);
void oracle_pack_rhs(int type, void* blockB, const void* rhs, long stride, long depth, long cols, int order, int conj);

/* ---- parallelize_gemm's partition (Parallelizer.h:85-157): returns thread count, fills per-thread slabs ---- */
int oracle_parallel_partition(int type, long rows, long cols, long depth, int max_threads, int transpose,
                              long* col0, long* ncols, long* row0, long* nrows /* each [max_threads] */);

/* ---- BLAS entry points with blas/level3_impl.h:12-76 semantics (sequential blocked path) ------------------ */
typedef int (*oracle_xerbla_fn)(const char* name, int* info, int len);
void oracle_set_xerbla(oracle_xerbla_fn fn); /* NULL -> default printer like blas/xerbla.cpp:15-19 */
int oracle_sgemm_(const char* ta, const char* tb, const int* m, const int* n, const int* k, const float* alpha,
                  const float* a, const int* lda, const float* b, const int* ldb, const float* beta, float* c,
                  const int* ldc);
int oracle_dgemm_(const char* ta, const char* tb, const int* m, const int* n, const int* k, const double* alpha,
                  const double* a, const int* lda, const double* b, const int* ldb, const double* beta, double* c,
                  const int* ldc);
int oracle_cgemm_(const char* ta, const char* tb, const int* m, const int* n, const int* k, const float* alpha,
                  const float* a, const int* lda, const float* b, const int* ldb, const float* beta, float* c,
                  const int* ldc);
int oracle_zgemm_(const char* ta, const char* tb, const int* m, const int* n, const int* k, const double* alpha,
                  const double* a, const int* lda, const double* b, const int* ldb, const double* beta, double* c,
                  const int* ldc);
/* threaded variant = parallelize_gemm + the OpenMP branch's column slabs (GeneralMatrixMatrix.h:83-152) */
int oracle_gemm_omp(int type, char ta, char tb, int m, int n, int k, const void* alpha, const void* a, int lda,
                    const void* b, int ldb, const void* beta, void* c, int ldc, int threads);

/* ---- rank-k updates, blas/level3_impl.h:357-433 (syrk), :564-627 (herk) -- oracle/rankk_port.c ---------------- */
int oracle_ssyrk_(const char* uplo, const char* trans, const int* n, const int* k, const float* alpha, const float* a, const int* lda, const float* beta, float* c, const int* ldc);
int oracle_dsyrk_(const char* uplo, const char* trans, const int* n, const int* k, const double* alpha, const double* a, const int* lda, const double* beta, double* c, const int* ldc);
int oracle_csyrk_(const char* uplo, const char* trans, const int* n, const int* k, const float* alpha, const float* a, const int* lda, const float* beta, float* c, const int* ldc);
int oracle_zsyrk_(const char* uplo, const char* trans, const int* n, const int* k, const double* alpha, const double* a, const int* lda, const double* beta, double* c, const int* ldc);
int oracle_cherk_(const char* uplo, const char* trans, const int* n, const int* k, const float* alpha, const float* a, const int* lda, const float* beta, float* c, const int* ldc);
int oracle_zherk_(const char* uplo, const char* trans, const int* n, const int* k, const double* alpha, const double* a, const int* lda, const double* beta, double* c, const int* ldc);
/* ---- trsm / trmm / symm / hemm / syr2k / her2k, blas/level3_impl.h:78-355,437-562,631-700 -- oracle/level3_port.c ---- */
#define ORACLE_DECL_TRI(NAME, R) int NAME(const char* side, const char* uplo, const char* opa, const char* diag, const int* m, const int* n, const R* alpha, const R* a, const int* lda, R* b, const int* ldb);
ORACLE_DECL_TRI(oracle_strsm_, float) ORACLE_DECL_TRI(oracle_dtrsm_, double) ORACLE_DECL_TRI(oracle_ctrsm_, float) ORACLE_DECL_TRI(oracle_ztrsm_, double)
ORACLE_DECL_TRI(oracle_strmm_, float) ORACLE_DECL_TRI(oracle_dtrmm_, double) ORACLE_DECL_TRI(oracle_ctrmm_, float) ORACLE_DECL_TRI(oracle_ztrmm_, double)
#define ORACLE_DECL_ABC(NAME, R) int NAME(const char* c1, const char* c2, const int* d1, const int* d2, const R* alpha, const R* a, const int* lda, const R* b, const int* ldb, const R* beta, R* c, const int* ldc);
ORACLE_DECL_ABC(oracle_ssymm_, float) ORACLE_DECL_ABC(oracle_dsymm_, double) ORACLE_DECL_ABC(oracle_csymm_, float) ORACLE_DECL_ABC(oracle_zsymm_, double)
ORACLE_DECL_ABC(oracle_chemm_, float) ORACLE_DECL_ABC(oracle_zhemm_, double)
ORACLE_DECL_ABC(oracle_ssyr2k_, float) ORACLE_DECL_ABC(oracle_dsyr2k_, double) ORACLE_DECL_ABC(oracle_csyr2k_, float) ORACLE_DECL_ABC(oracle_zsyr2k_, double)
ORACLE_DECL_ABC(oracle_cher2k_, float) ORACLE_DECL_ABC(oracle_zher2k_, double)
/* ---- potrf / getrf, lapack/cholesky.cpp:14-38, lapack/lu.cpp:14-42 -- oracle/lapack_port.c ---------------------- */
int oracle_spotrf_(const char* uplo, const int* n, float* a, const int* lda, int* info);
int oracle_dpotrf_(const char* uplo, const int* n, double* a, const int* lda, int* info);
int oracle_cpotrf_(const char* uplo, const int* n, float* a, const int* lda, int* info);
int oracle_zpotrf_(const char* uplo, const int* n, double* a, const int* lda, int* info);
int oracle_sgetrf_(const int* m, const int* n, float* a, const int* lda, int* ipiv, int* info);
int oracle_dgetrf_(const int* m, const int* n, double* a, const int* lda, int* ipiv, int* info);
int oracle_cgetrf_(const int* m, const int* n, float* a, const int* lda, int* ipiv, int* info);
int oracle_zgetrf_(const int* m, const int* n, double* a, const int* lda, int* ipiv, int* info);
void oracle_xerbla_expect(const char* name6, int infot);
int oracle_xerbla_result(void);

/* ---- high-precision reference (long double accumulate), column-at-a-time like DMMCH --------------------- */
/* Computes rows listed in row_idx[nrows] (or all rows if row_idx==NULL) of C_ref = alpha*op(A)*op(B)+beta*C
 * into out (nrows x n, column-major, ld = nrows; ALWAYS double / double-complex, also for s and c) and the
 * gauge G = |alpha|sum|a||b|+|beta||c| (double, nrows x n). */
void oracle_hp_gemm(int type, char ta, char tb, int m, int n, int k, const void* alpha, const void* a, int lda,
                    const void* b, int ldb, const void* beta, const void* c, int ldc, const int* row_idx, int nrows,
                    double* out, double* gauge);

/* ---- xBLAT3 restatement (blas/testing/{s,d,c,z}blat3.f) ------------------------------------------------- */
typedef int (*oracle_gemm_fn)(const char*, const char*, const int*, const int*, const int*, const void*, const void*,
                              const int*, const void*, const int*, const void*, void*, const int*);
typedef struct {
  int ncalls;        /* NC */
  double errmax;     /* ERRMAX (test ratio) */
  int fatal;         /* 1 = FATAL (argument changed, error exit on valid call, or < half accurate) */
  int bad_param;     /* ISAME index (1..13) that changed, 0 if none */
  char msg[200];
} oracle_blat3_report;
/* xCHK1 for GEMM: dims from .dat (0 1 2 3 5 9), NMAX=65, ld=dim+1, 3 alphas x 3 betas, N/T/C x N/T/C */
void oracle_blat3_chk1(int type, oracle_gemm_fn gemm, oracle_blat3_report* rep);
/* xCHKE GEMM block: returns number of failed error-exit cases (0 = pass).  set_xerbla installs the tester's
 * XERBLA replacement into the library under test (may be NULL if the library resolves xerbla_ by symbol). */
typedef void (*oracle_install_xerbla_fn)(oracle_xerbla_fn);
int oracle_blat3_chke(int type, oracle_gemm_fn gemm, oracle_install_xerbla_fn install, char* log, int loglen);
/* generators, exposed for tests */
void oracle_blat3_reset(void);
double oracle_blat3_dbeg(void);
void oracle_blat3_zbeg(double* re, double* im);

#ifdef __cplusplus
}
#endif
#endif
