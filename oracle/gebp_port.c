/* oracle/gebp_port.c -- CPU ORACLE, TEST INFRASTRUCTURE ONLY (see oracle.h for the contract and pin status).
 *
 * Plain-C restatement of Eigen's blocked GEMM for the four BLAS scalar types:
 *   cache model + blocking heuristic   Eigen/src/Core/products/GeneralBlockPanelKernel.h:39-78, 92-308
 *   parallelize_gemm partition         Eigen/src/Core/products/Parallelizer.h:85-157
 *   pack / gebp / run / blas gemm      see gebp_impl.h
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

#define ORACLE_SIMD_BYTES 32 /* the reference oracle build is -mavx2 -mfma */

/* CacheSizes defaults when CPUID gives nothing (GeneralBlockPanelKernel.h:39-60); tests overwrite them with the
 * values the reference detected on the running host so that kc/mc/nc, and therefore rounding, coincide. */
static long g_l1 = 32 * 1024, g_l2 = 256 * 1024, g_l3 = 2048 * 1024;
void oracle_set_cache_sizes(long l1, long l2, long l3) { g_l1 = l1; g_l2 = l2; g_l3 = l3; }
void oracle_get_cache_sizes(long* l1, long* l2, long* l3) { *l1 = g_l1; *l2 = g_l2; *l3 = g_l3; }

static const int k_scalar_bytes[4] = {4, 8, 8, 16};
void oracle_gebp_traits(int type, int* mr, int* nr, int* lhs_progress) {
  const int P = ORACLE_SIMD_BYTES / k_scalar_bytes[type];
  *lhs_progress = P;
  *mr = type < 2 ? 3 * P : P; /* :369-380 (FMA => 3 packets), :618-619 (complex: 1 packet) */
  *nr = 4;
}

static long lmin(long a, long b) { return a < b ? a : b; }
static long lmax(long a, long b) { return a > b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }

/* evaluateProductBlockingSizesHeuristic<S,S,KcFactor=1> (GeneralBlockPanelKernel.h:92-259) */
void oracle_blocking_sizes(int type, long* pk, long* pm, long* pn, int num_threads) {
  long k = *pk, m = *pm, n = *pn;
  int mr, nr, lp;
  oracle_gebp_traits(type, &mr, &nr, &lp);
  const long sz = k_scalar_bytes[type];
  const long l1 = g_l1, l2 = g_l2, l3 = g_l3;
  if (num_threads > 1) { /* :104-149 */
    const long kdiv = mr * sz + nr * sz, ksub = mr * nr * sz, kr = 8;
    const long k_cache = lmin((l1 - ksub) / kdiv, 320);
    if (k_cache < k) k = k_cache - (k_cache % kr);
    const long n_cache = (l2 - l1) / (nr * sz * k);
    const long n_per_thread = (n + num_threads - 1) / num_threads;
    if (n_cache <= n_per_thread) n = n_cache - (n_cache % nr);
    else n = lmin(n, (n_per_thread + nr - 1) - ((n_per_thread + nr - 1) % nr));
    if (l3 > l2) {
      const long m_cache = (l3 - l2) / (sz * k * num_threads);
      const long m_per_thread = (m + num_threads - 1) / num_threads;
      if (m_cache < m_per_thread && m_cache >= mr) m = m_cache - (m_cache % mr);
      else m = lmin(m, (m_per_thread + mr - 1) - ((m_per_thread + mr - 1) % mr));
    }
  } else { /* :150-258 */
    if (lmax(k, lmax(m, n)) < 48) return;
    const long k_peeling = 8, k_div = mr * sz + nr * sz, k_sub = mr * nr * sz;
    const long max_kc = lmax(((l1 - k_sub) / k_div) & ~(k_peeling - 1), 1);
    const long old_k = k;
    if (k > max_kc)
      k = (k % max_kc) == 0 ? max_kc : max_kc - k_peeling * ((max_kc - 1 - (k % max_kc)) / (k_peeling * (k / max_kc + 1)));
    const long actual_l2 = 1572864;
    long max_nc;
    const long lhs_bytes = m * k * sz;
    const long remaining_l1 = l1 - k_sub - lhs_bytes;
    if (remaining_l1 >= nr * sz * k) max_nc = remaining_l1 / (k * sz);
    else max_nc = (3 * actual_l2) / (2 * 2 * max_kc * sz);
    const long nc = lmin(actual_l2 / (2 * k * sz), max_nc) & ~(long)(nr - 1);
    if (n > nc) {
      n = (n % nc) == 0 ? nc : (nc - nr * ((nc - (n % nc)) / (nr * (n / nc + 1))));
    } else if (old_k == k) {
      const long problem_size = k * n * sz;
      long actual_lm = actual_l2, max_mc = m;
      if (problem_size <= 1024) actual_lm = l1;
      else if (l3 != 0 && problem_size <= 32768) { actual_lm = l2; max_mc = lmin(576, max_mc); }
      long mc = lmin(actual_lm / (3 * k * sz), max_mc);
      if (mc > mr) mc -= mc % mr;
      else if (mc == 0) { *pk = k; *pm = m; *pn = n; return; }
      m = (m % mc) == 0 ? mc : (mc - mr * ((mc - (m % mc)) / (mr * (m / mc + 1))));
    }
  }
  *pk = k; *pm = m; *pn = n;
}

/* parallelize_gemm (Parallelizer.h:85-157): thread count heuristic (:108-118) and the slab split (:134-155).
 * transpose swaps the roles of rows/cols (row-major destination). */
int oracle_parallel_partition(int type, long rows, long cols, long depth, int max_threads, int transpose, long* col0,
                              long* ncols, long* row0, long* nrows) {
  int mr, nr, lp;
  oracle_gebp_traits(type, &mr, &nr, &lp);
  const long size = transpose ? rows : cols;
  long pb_max_threads = lmax(1, size / nr);
  const double work = (double)rows * (double)cols * (double)depth;
  pb_max_threads = lmax(1, lmin(pb_max_threads, (long)(work / 50000.0)));
  const long threads = lmin(max_threads, pb_max_threads);
  if (threads <= 1) {
    col0[0] = 0; ncols[0] = cols; row0[0] = 0; nrows[0] = rows;
    return 1;
  }
  if (transpose) { const long t = rows; rows = cols; cols = t; }
  const long blockCols = (cols / threads) & ~(long)0x3;
  long blockRows = rows / threads;
  blockRows = (blockRows / mr) * mr;
  for (long i = 0; i < threads; ++i) {
    row0[i] = i * blockRows;
    nrows[i] = (i + 1 == threads) ? rows - row0[i] : blockRows;
    col0[i] = i * blockCols;
    ncols[i] = (i + 1 == threads) ? cols - col0[i] : blockCols;
  }
  return (int)threads;
}

/* OP() of blas/common.h:39-42 */
static int oracle_op(char x) {
  return (x == 'N' || x == 'n') ? ORACLE_NOTR : (x == 'T' || x == 't') ? ORACLE_TR : (x == 'C' || x == 'c') ? ORACLE_ADJ : ORACLE_INVALID;
}

/* xerbla_ (blas/xerbla.cpp:15-19) with a test hook */
static oracle_xerbla_fn g_xerbla = NULL;
void oracle_set_xerbla(oracle_xerbla_fn fn) { g_xerbla = fn; }
int xerbla_(const char* name, int* info, int len); /* blat3_port.c: the tester's XERBLA, prints like blas/xerbla.cpp when unarmed */
static int oracle_call_xerbla(const char* name, int* info) {
  if (g_xerbla) return g_xerbla(name, info, 6);
  return xerbla_(name, info, 6);
}

int oracle_call_xerbla_public(const char* name, int* info) { return oracle_call_xerbla(name, info); }

#define R float
#define NC 1
#define SFX s
#define FMA fmaf
#define TYPE_CODE ORACLE_S
#include "gebp_impl.h"
#undef R
#undef NC
#undef SFX
#undef FMA
#undef TYPE_CODE

#define R double
#define NC 1
#define SFX d
#define FMA fma
#define TYPE_CODE ORACLE_D
#include "gebp_impl.h"
#undef R
#undef NC
#undef SFX
#undef FMA
#undef TYPE_CODE

#define R float
#define NC 2
#define SFX c
#define FMA fmaf
#define TYPE_CODE ORACLE_C
#include "gebp_impl.h"
#undef R
#undef NC
#undef SFX
#undef FMA
#undef TYPE_CODE

#define R double
#define NC 2
#define SFX z
#define FMA fma
#define TYPE_CODE ORACLE_Z
#include "gebp_impl.h"
#undef R
#undef NC
#undef SFX
#undef FMA
#undef TYPE_CODE

void oracle_pack_lhs(int type, void* blockA, const void* lhs, long stride, long depth, long rows, int order, int conj) {
  switch (type) {
    case ORACLE_S: pack_lhs_s((float*)blockA, (const float*)lhs, stride, order, conj, depth, rows); break;
    case ORACLE_D: pack_lhs_d((double*)blockA, (const double*)lhs, stride, order, conj, depth, rows); break;
    case ORACLE_C: pack_lhs_c((float*)blockA, (const float*)lhs, stride, order, conj, depth, rows); break;
    default: pack_lhs_z((double*)blockA, (const double*)lhs, stride, order, conj, depth, rows); break;
  }
}
void oracle_pack_rhs(int type, void* blockB, const void* rhs, long stride, long depth, long cols, int order, int conj) {
  switch (type) {
    case ORACLE_S: pack_rhs_s((float*)blockB, (const float*)rhs, stride, order, conj, depth, cols); break;
    case ORACLE_D: pack_rhs_d((double*)blockB, (const double*)rhs, stride, order, conj, depth, cols); break;
    case ORACLE_C: pack_rhs_c((float*)blockB, (const float*)rhs, stride, order, conj, depth, cols); break;
    default: pack_rhs_z((double*)blockB, (const double*)rhs, stride, order, conj, depth, cols); break;
  }
}

int oracle_sgemm_(const char* ta, const char* tb, const int* m, const int* n, const int* k, const float* alpha,
                  const float* a, const int* lda, const float* b, const int* ldb, const float* beta, float* c,
                  const int* ldc) {
  return blas_gemm_s("SGEMM ", ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, 1);
}
int oracle_dgemm_(const char* ta, const char* tb, const int* m, const int* n, const int* k, const double* alpha,
                  const double* a, const int* lda, const double* b, const int* ldb, const double* beta, double* c,
                  const int* ldc) {
  return blas_gemm_d("DGEMM ", ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, 1);
}
int oracle_cgemm_(const char* ta, const char* tb, const int* m, const int* n, const int* k, const float* alpha,
                  const float* a, const int* lda, const float* b, const int* ldb, const float* beta, float* c,
                  const int* ldc) {
  return blas_gemm_c("CGEMM ", ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, 1);
}
int oracle_zgemm_(const char* ta, const char* tb, const int* m, const int* n, const int* k, const double* alpha,
                  const double* a, const int* lda, const double* b, const int* ldb, const double* beta, double* c,
                  const int* ldc) {
  return blas_gemm_z("ZGEMM ", ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, 1);
}
int oracle_gemm_omp(int type, char ta, char tb, int m, int n, int k, const void* alpha, const void* a, int lda,
                    const void* b, int ldb, const void* beta, void* c, int ldc, int threads) {
  switch (type) {
    case ORACLE_S: return blas_gemm_s("SGEMM ", &ta, &tb, &m, &n, &k, (const float*)alpha, (const float*)a, &lda, (const float*)b, &ldb, (const float*)beta, (float*)c, &ldc, threads);
    case ORACLE_D: return blas_gemm_d("DGEMM ", &ta, &tb, &m, &n, &k, (const double*)alpha, (const double*)a, &lda, (const double*)b, &ldb, (const double*)beta, (double*)c, &ldc, threads);
    case ORACLE_C: return blas_gemm_c("CGEMM ", &ta, &tb, &m, &n, &k, (const float*)alpha, (const float*)a, &lda, (const float*)b, &ldb, (const float*)beta, (float*)c, &ldc, threads);
    default: return blas_gemm_z("ZGEMM ", &ta, &tb, &m, &n, &k, (const double*)alpha, (const double*)a, &lda, (const double*)b, &ldb, (const double*)beta, (double*)c, &ldc, threads);
  }
}
