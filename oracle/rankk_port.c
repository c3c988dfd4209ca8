/* oracle/rankk_port.c -- CPU ORACLE, TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Plain-C restatement of the reference's rank-k updates:
 *   EIGEN_BLAS_FUNC(syrk)   blas/level3_impl.h:357-433   C.tri = alpha*op(A)*op(A)^T + beta*C.tri   (s, d, c, z)
 *   EIGEN_BLAS_FUNC(herk)   blas/level3_impl.h:564-627   C.tri = alpha*op(A)*op(A)^H + beta*C.tri   (c, z; real alpha, beta)
 * Order of operations as in the reference: argument checks -> xerbla_, beta pre-pass over the referenced triangle
 * (herk: strict triangle scaled, diagonal real part scaled and imaginary part zeroed), early returns, then the product
 * accumulated per kc block (general_matrix_matrix_triangular_product, GeneralMatrixMatrixTriangular.h:36-135: the
 * same pack + gebp blocks as GEMM, i.e. one FMA chain per kc block and element, then one update of C).
 */
#include <math.h>
#include <stdio.h>
#include <string.h>
#include "oracle.h"

int oracle_call_xerbla_public(const char* name, int* info); /* gebp_port.c */

static int op_code(char x) {
  return (x == 'N' || x == 'n') ? ORACLE_NOTR : (x == 'T' || x == 't') ? ORACLE_TR : (x == 'C' || x == 'c') ? ORACLE_ADJ : ORACLE_INVALID;
}
static int imax2(int a, int b) { return a > b ? a : b; }

#define RANKK_IMPL(SFX, R, CPLX, FMA, TYPE_CODE, NAME_SYRK, NAME_HERK)                                                  \
  static int rankk_##SFX(int herk, const char* uplo, const char* op, const int* pn, const int* pk, const R* alpha,        \
                         const R* a, const int* plda, const R* beta, R* c, const int* pldc) {                             \
    const int NCc = CPLX ? 2 : 1;                                                                                         \
    const int up = (*uplo == 'U' || *uplo == 'u') ? 1 : (*uplo == 'L' || *uplo == 'l') ? 0 : -1;                         \
    const int o = op_code(*op);                                                                                           \
    int info = 0;                                                                                                         \
    if (up < 0) info = 1;                                                                                                 \
    else if (o == ORACLE_INVALID || (!herk && CPLX && o == ORACLE_ADJ) || (herk && o == ORACLE_TR)) info = 2;             \
    else if (*pn < 0) info = 3;                                                                                           \
    else if (*pk < 0) info = 4;                                                                                           \
    else if (*plda < imax2(1, o == ORACLE_NOTR ? *pn : *pk)) info = 7;                                                    \
    else if (*pldc < imax2(1, *pn)) info = 10;                                                                            \
    if (info) return oracle_call_xerbla_public(herk ? NAME_HERK : NAME_SYRK, &info);                                      \
    const long n = *pn, k = *pk, lda = *plda, ldc = *pldc;                                                                \
    const R ar = alpha[0], ai = (CPLX && !herk) ? alpha[1] : (R)0, br = beta[0], bi = (CPLX && !herk) ? beta[1] : (R)0;   \
    const int beta_one = (br == (R)1 && bi == (R)0), beta_zero = (br == (R)0 && bi == (R)0);                              \
    if (!beta_one) { /* beta pre-pass, :389-395 / :598-611 */                                                             \
      for (long j = 0; j < n; ++j)                                                                                        \
        for (long i = (up ? 0 : j); i < (up ? j + 1 : n); ++i) {                                                          \
          R* z = c + NCc * (i + j * ldc);                                                                                 \
          if (beta_zero) { z[0] = 0; if (CPLX) z[NCc - 1] = 0; }                                                          \
          else if (!CPLX) z[0] = z[0] * br;                                                                               \
          else if (herk) { z[0] = z[0] * br; z[1] = (i == j) ? (R)0 : z[1] * br; }                                        \
          else { const R re = z[0] * br - z[1] * bi, im = z[0] * bi + z[1] * br; z[0] = re; z[1] = im; }                  \
        }                                                                                                                 \
    }                                                                                                                     \
    if (n == 0 || k == 0) return 0;                                                                                       \
    if (herk && ar == (R)0) return 0; /* :617 */                                                                          \
    long kc = k, mc = n, nc = n;                                                                                          \
    oracle_blocking_sizes(TYPE_CODE, &kc, &mc, &nc, 1);                                                                   \
    for (long k0 = 0; k0 < k; k0 += kc) {                                                                                 \
      const long k1 = k0 + kc < k ? k0 + kc : k;                                                                          \
      for (long j = 0; j < n; ++j)                                                                                        \
        for (long i = (up ? 0 : j); i < (up ? j + 1 : n); ++i) {                                                          \
          R f0 = 0, f1 = 0, s0 = 0, s1 = 0;                                                                               \
          for (long p = k0; p < k1; ++p) {                                                                                \
            const R* x = a + NCc * (o == ORACLE_NOTR ? i + p * lda : p + i * lda); /* op(A)(i,p) */                       \
            const R* y = a + NCc * (o == ORACLE_NOTR ? j + p * lda : p + j * lda); /* op(A)(j,p) */                       \
            if (!CPLX) { f0 = FMA(x[0], y[0], f0); }                                                                      \
            else {                                                                                                        \
              R xr = x[0], xi = x[NCc - 1], yr = y[0], yi = y[NCc - 1];                                                   \
              if (herk) { if (o == ORACLE_NOTR) yi = -yi; else xi = -xi; } /* A*A^H or A^H*A */                           \
              f0 = FMA(xr, yr, f0); f1 = FMA(xi, yr, f1); s0 = FMA(xr, yi, s0); s1 = FMA(xi, yi, s1);                     \
            }                                                                                                             \
          }                                                                                                               \
          R* z = c + NCc * (i + j * ldc);                                                                                 \
          if (!CPLX) z[0] = FMA(f0, ar, z[0]);                                                                            \
          else {                                                                                                          \
            const R t0 = f0 - s1, t1 = f1 + s0;                                                                           \
            z[0] += t0 * ar - t1 * ai; z[NCc - 1] += t0 * ai + t1 * ar;                                                   \
          }                                                                                                               \
        }                                                                                                                 \
    }                                                                                                                     \
    if (herk) for (long j = 0; j < n; ++j) c[NCc * (j + j * ldc) + NCc - 1] = 0; /* :621 diagonal().imag().setZero() */   \
    return 0;                                                                                                             \
  }

RANKK_IMPL(s, float, 0, fmaf, ORACLE_S, "SSYRK ", "")
RANKK_IMPL(d, double, 0, fma, ORACLE_D, "DSYRK ", "")
RANKK_IMPL(c, float, 1, fmaf, ORACLE_C, "CSYRK ", "CHERK ")
RANKK_IMPL(z, double, 1, fma, ORACLE_Z, "ZSYRK ", "ZHERK ")

int oracle_ssyrk_(const char* u, const char* t, const int* n, const int* k, const float* al, const float* a, const int* lda, const float* be, float* c, const int* ldc) { return rankk_s(0, u, t, n, k, al, a, lda, be, c, ldc); }
int oracle_dsyrk_(const char* u, const char* t, const int* n, const int* k, const double* al, const double* a, const int* lda, const double* be, double* c, const int* ldc) { return rankk_d(0, u, t, n, k, al, a, lda, be, c, ldc); }
int oracle_csyrk_(const char* u, const char* t, const int* n, const int* k, const float* al, const float* a, const int* lda, const float* be, float* c, const int* ldc) { return rankk_c(0, u, t, n, k, al, a, lda, be, c, ldc); }
int oracle_zsyrk_(const char* u, const char* t, const int* n, const int* k, const double* al, const double* a, const int* lda, const double* be, double* c, const int* ldc) { return rankk_z(0, u, t, n, k, al, a, lda, be, c, ldc); }
int oracle_cherk_(const char* u, const char* t, const int* n, const int* k, const float* al, const float* a, const int* lda, const float* be, float* c, const int* ldc) { return rankk_c(1, u, t, n, k, al, a, lda, be, c, ldc); }
int oracle_zherk_(const char* u, const char* t, const int* n, const int* k, const double* al, const double* a, const int* lda, const double* be, double* c, const int* ldc) { return rankk_z(1, u, t, n, k, al, a, lda, be, c, ldc); }
