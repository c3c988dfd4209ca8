/* oracle/level3_port.c -- CPU ORACLE, TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Plain-C (C99 complex) restatement of the reference's remaining level-3 routines:
 *   EIGEN_BLAS_FUNC(trsm)    blas/level3_impl.h:78-178     B := alpha * inv(op(A)) * B  |  alpha * B * inv(op(A))
 *   EIGEN_BLAS_FUNC(trmm)    blas/level3_impl.h:183-284    B := alpha * op(A) * B       |  alpha * B * op(A)
 *   EIGEN_BLAS_FUNC(symm)    blas/level3_impl.h:287-355    C := alpha * A * B + beta * C |  alpha * B * A + beta * C
 *   EIGEN_BLAS_FUNC(hemm)    blas/level3_impl.h:505-562
 *   EIGEN_BLAS_FUNC(syr2k)   blas/level3_impl.h:437-503    C.tri := alpha*op(A)op(B)^T + alpha*op(B)op(A)^T + beta*C.tri
 *   EIGEN_BLAS_FUNC(her2k)   blas/level3_impl.h:631-700
 * Order of operations as in the reference: argument checks -> xerbla_ with the reference's info codes and names, quick
 * returns and RETURN VALUES as written there (trmm returns 1, symm/hemm return 1 on an empty result, syr2k returns 1
 * when k == 0, her2k always returns 1), beta pre-pass, product.  trsm: substitution with the reciprocal of the
 * diagonal (TriangularSolverMatrix.h:118-121) and alpha applied AFTER the solve (level3_impl.h:174-175).  trmm: B is
 * copied, zeroed, and the product accumulated into it (level3_impl.h:267-281).  The summation order inside each
 * element is one FMA-free chain over the inner index; parity with the compiled reference is pinned within a few
 * gauge units by tests/test_oracle_pin_level3.py.
 */
#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

int oracle_call_xerbla_public(const char* name, int* info); /* gebp_port.c */

static int op_code(char x) {
  return (x == 'N' || x == 'n') ? ORACLE_NOTR : (x == 'T' || x == 't') ? ORACLE_TR : (x == 'C' || x == 'c') ? ORACLE_ADJ : ORACLE_INVALID;
}
static int side_code(char x) { return (x == 'L' || x == 'l') ? 1 : (x == 'R' || x == 'r') ? 0 : -1; }
static int uplo_code(char x) { return (x == 'U' || x == 'u') ? 1 : (x == 'L' || x == 'l') ? 0 : -1; }
static int diag_code(char x) { return (x == 'U' || x == 'u') ? 1 : (x == 'N' || x == 'n') ? 0 : -1; }
static int imax2(int a, int b) { return a > b ? a : b; }

#define CJ_float(x) (x)
#define CJ_double(x) (x)
#define CJ_cfloat(x) conjf(x)
#define CJ_cdouble(x) conj(x)
typedef float _Complex cfloat;
typedef double _Complex cdouble;

/* element (i, j) of op(A) for a triangular A (zero outside the stored triangle, one on a unit diagonal) */
#define L3_IMPL(SFX, T, R, CPLX, CJ, NM)                                                                                      \
  static T tri_at_##SFX(const T* a, long lda, int up, int o, int unit, long i, long j) {                                      \
    const long r = (o == ORACLE_NOTR) ? i : j, c = (o == ORACLE_NOTR) ? j : i;                                                \
    if (r == c) { if (unit) return (T)1; }                                                                                    \
    else if (up ? r > c : r < c) return (T)0;                                                                                 \
    const T v = a[r + c * lda];                                                                                               \
    return (o == ORACLE_ADJ) ? CJ(v) : v;                                                                                     \
  }                                                                                                                           \
  static int tri_args_##SFX(const char* name, const char* side, const char* uplo, const char* opa, const char* diag,          \
                            const int* m, const int* n, const int* lda, const int* ldb) {                                     \
    int info = 0;                                                                                                             \
    if (side_code(*side) < 0) info = 1;                                                                                       \
    else if (uplo_code(*uplo) < 0) info = 2;                                                                                  \
    else if (op_code(*opa) == ORACLE_INVALID) info = 3;                                                                       \
    else if (diag_code(*diag) < 0) info = 4;                                                                                  \
    else if (*m < 0) info = 5;                                                                                                \
    else if (*n < 0) info = 6;                                                                                                \
    else if (*lda < imax2(1, side_code(*side) ? *m : *n)) info = 9;                                                           \
    else if (*ldb < imax2(1, *m)) info = 11;                                                                                  \
    if (info) { oracle_call_xerbla_public(name, &info); return 1; }                                                           \
    return 0;                                                                                                                 \
  }                                                                                                                           \
  int oracle_##SFX##trsm_(const char* side, const char* uplo, const char* opa, const char* diag, const int* pm, const int* pn,\
                          const R* palpha, const R* pa, const int* plda, R* pb, const int* pldb) {                            \
    if (tri_args_##SFX(NM "TRSM ", side, uplo, opa, diag, pm, pn, plda, pldb)) return 0;                                      \
    const long m = *pm, n = *pn, lda = *plda, ldb = *pldb;                                                                    \
    if (m == 0 || n == 0) return 0;                                                                                           \
    const int left = side_code(*side), up = uplo_code(*uplo), o = op_code(*opa), unit = diag_code(*diag);                     \
    const T* a = (const T*)pa; T* b = (T*)pb; const T alpha = *(const T*)palpha;                                              \
    const long na = left ? m : n;                                                                                             \
    const int t_lower = (up == 0) == (o == ORACLE_NOTR); /* op(A) is lower triangular */                                      \
    if (left) {                                                                                                               \
      for (long j = 0; j < n; ++j) {                                                                                          \
        T* x = b + j * ldb;                                                                                                   \
        for (long s = 0; s < na; ++s) {                                                                                       \
          const long i = t_lower ? s : na - 1 - s;                                                                            \
          T acc = x[i];                                                                                                       \
          for (long p = (t_lower ? 0 : i + 1); p < (t_lower ? i : na); ++p) acc -= tri_at_##SFX(a, lda, up, o, unit, i, p) * x[p]; \
          x[i] = unit ? acc : acc * ((T)1 / tri_at_##SFX(a, lda, up, o, unit, i, i));                                         \
        }                                                                                                                     \
      }                                                                                                                       \
    } else { /* X op(A) = B: column j of X depends on the columns already solved */                                           \
      for (long s = 0; s < na; ++s) {                                                                                         \
        const long j = t_lower ? na - 1 - s : s;                                                                              \
        const T d = unit ? (T)1 : (T)1 / tri_at_##SFX(a, lda, up, o, unit, j, j);                                             \
        for (long i = 0; i < m; ++i) {                                                                                        \
          T acc = b[i + j * ldb];                                                                                             \
          for (long p = (t_lower ? j + 1 : 0); p < (t_lower ? na : j); ++p) acc -= b[i + p * ldb] * tri_at_##SFX(a, lda, up, o, unit, p, j); \
          b[i + j * ldb] = unit ? acc : acc * d;                                                                              \
        }                                                                                                                     \
      }                                                                                                                       \
    }                                                                                                                         \
    if (alpha != (T)1)                                                                                                        \
      for (long j = 0; j < n; ++j) for (long i = 0; i < m; ++i) b[i + j * ldb] *= alpha;                                      \
    return 0;                                                                                                                 \
  }                                                                                                                           \
  int oracle_##SFX##trmm_(const char* side, const char* uplo, const char* opa, const char* diag, const int* pm, const int* pn,\
                          const R* palpha, const R* pa, const int* plda, R* pb, const int* pldb) {                            \
    if (tri_args_##SFX(NM "TRMM ", side, uplo, opa, diag, pm, pn, plda, pldb)) return 0;                                      \
    const long m = *pm, n = *pn, lda = *plda, ldb = *pldb;                                                                    \
    if (m == 0 || n == 0) return 1;                                                                                           \
    const int left = side_code(*side), up = uplo_code(*uplo), o = op_code(*opa), unit = diag_code(*diag);                     \
    const T* a = (const T*)pa; T* b = (T*)pb; const T alpha = *(const T*)palpha;                                              \
    T* tmp = (T*)malloc(sizeof(T) * (size_t)m * (size_t)n);                                                                   \
    for (long j = 0; j < n; ++j) for (long i = 0; i < m; ++i) { tmp[i + j * m] = b[i + j * ldb]; b[i + j * ldb] = (T)0; }     \
    for (long j = 0; j < n; ++j)                                                                                              \
      for (long i = 0; i < m; ++i) {                                                                                          \
        T acc = (T)0;                                                                                                         \
        if (left) for (long p = 0; p < m; ++p) acc += tri_at_##SFX(a, lda, up, o, unit, i, p) * tmp[p + j * m];               \
        else for (long p = 0; p < n; ++p) acc += tmp[i + p * m] * tri_at_##SFX(a, lda, up, o, unit, p, j);                    \
        b[i + j * ldb] += alpha * acc;                                                                                        \
      }                                                                                                                       \
    free(tmp);                                                                                                                \
    return 1;                                                                                                                 \
  }                                                                                                                           \
  /* element (i, j) of the symmetric / Hermitian matrix stored in one triangle */                                             \
  static T sym_at_##SFX(const T* a, long lda, int up, int herm, long i, long j) {                                             \
    if (i == j) return herm ? (T)creal((cdouble)a[i + i * lda]) : a[i + i * lda];                                             \
    const int stored = up ? i < j : i > j;                                                                                    \
    if (stored) return a[i + j * lda];                                                                                        \
    return herm ? CJ(a[j + i * lda]) : a[j + i * lda];                                                                        \
  }                                                                                                                           \
  static int symm_##SFX(int herm, const char* name, const char* side, const char* uplo, const int* pm, const int* pn,         \
                        const R* palpha, const R* pa, const int* plda, const R* pb, const int* pldb, const R* pbeta, R* pc,  \
                        const int* pldc) {                                                                                    \
    int info = 0;                                                                                                             \
    if (side_code(*side) < 0) info = 1;                                                                                       \
    else if (uplo_code(*uplo) < 0) info = 2;                                                                                  \
    else if (*pm < 0) info = 3;                                                                                               \
    else if (*pn < 0) info = 4;                                                                                               \
    else if (*plda < imax2(1, side_code(*side) ? *pm : *pn)) info = 7;                                                        \
    else if (*pldb < imax2(1, *pm)) info = 9;                                                                                 \
    else if (*pldc < imax2(1, *pm)) info = 12;                                                                                \
    if (info) return oracle_call_xerbla_public(name, &info);                                                                  \
    const long m = *pm, n = *pn, lda = *plda, ldb = *pldb, ldc = *pldc;                                                       \
    const int left = side_code(*side), up = uplo_code(*uplo);                                                                 \
    const T* a = (const T*)pa; const T* b = (const T*)pb; T* c = (T*)pc;                                                      \
    const T alpha = *(const T*)palpha, beta = *(const T*)pbeta;                                                               \
    if (beta != (T)1)                                                                                                         \
      for (long j = 0; j < n; ++j) for (long i = 0; i < m; ++i) c[i + j * ldc] = (beta == (T)0) ? (T)0 : c[i + j * ldc] * beta; \
    if (m == 0 || n == 0) return 1;                                                                                           \
    for (long j = 0; j < n; ++j)                                                                                              \
      for (long i = 0; i < m; ++i) {                                                                                          \
        T acc = (T)0;                                                                                                         \
        if (left) for (long p = 0; p < m; ++p) acc += sym_at_##SFX(a, lda, up, herm, i, p) * b[p + j * ldb];                  \
        else for (long p = 0; p < n; ++p) acc += b[i + p * ldb] * sym_at_##SFX(a, lda, up, herm, p, j);                       \
        c[i + j * ldc] += alpha * acc;                                                                                        \
      }                                                                                                                       \
    return 0;                                                                                                                 \
  }                                                                                                                           \
  static int r2k_##SFX(int her, const char* name, const char* uplo, const char* op, const int* pn, const int* pk,             \
                       const R* palpha, const R* pa, const int* plda, const R* pb, const int* pldb, const R* pbeta, R* pc,   \
                       const int* pldc) {                                                                                     \
    const int up = uplo_code(*uplo), o = op_code(*op);                                                                        \
    int info = 0;                                                                                                             \
    if (up < 0) info = 1;                                                                                                     \
    else if (o == ORACLE_INVALID || (!her && CPLX && o == ORACLE_ADJ) || (her && o == ORACLE_TR)) info = 2;                   \
    else if (*pn < 0) info = 3;                                                                                               \
    else if (*pk < 0) info = 4;                                                                                               \
    else if (*plda < imax2(1, o == ORACLE_NOTR ? *pn : *pk)) info = 7;                                                        \
    else if (*pldb < imax2(1, o == ORACLE_NOTR ? *pn : *pk)) info = 9;                                                        \
    else if (*pldc < imax2(1, *pn)) info = 12;                                                                                \
    if (info) return oracle_call_xerbla_public(name, &info);                                                                  \
    const long n = *pn, k = *pk, lda = *plda, ldb = *pldb, ldc = *pldc;                                                       \
    const T* a = (const T*)pa; const T* b = (const T*)pb; T* c = (T*)pc;                                                      \
    const T alpha = *(const T*)palpha;                                                                                        \
    const T beta = her ? (T)pbeta[0] : *(const T*)pbeta; /* her2k: REAL beta (:638) */                                        \
    if (beta != (T)1) { /* :457-467 / :653-669 */                                                                             \
      for (long j = 0; j < n; ++j)                                                                                            \
        for (long i = (up ? 0 : j); i < (up ? j + 1 : n); ++i) {                                                              \
          T* z = c + i + j * ldc;                                                                                             \
          if (beta == (T)0) *z = (T)0;                                                                                        \
          else if (her && i == j) *z = (T)(creal((cdouble)*z) * creal((cdouble)beta));                                        \
          else *z = *z * beta;                                                                                                \
        }                                                                                                                     \
    } else if (her && k > 0 && alpha != (T)0) {                                                                               \
      for (long j = 0; j < n; ++j) c[j + j * ldc] = (T)creal((cdouble)c[j + j * ldc]);                                        \
    }                                                                                                                         \
    if (k == 0) return 1;                                                                                                     \
    const T alpha2 = her ? CJ(alpha) : alpha;                                                                                 \
    for (long j = 0; j < n; ++j)                                                                                              \
      for (long i = (up ? 0 : j); i < (up ? j + 1 : n); ++i) {                                                                \
        T s1 = (T)0, s2 = (T)0;                                                                                               \
        for (long p = 0; p < k; ++p) {                                                                                        \
          if (o == ORACLE_NOTR) { /* A B^T|^H and B A^T|^H */                                                                 \
            const T ai = a[i + p * lda], aj = a[j + p * lda], bi = b[i + p * ldb], bj = b[j + p * ldb];                       \
            s1 += ai * (her ? CJ(bj) : bj); s2 += bi * (her ? CJ(aj) : aj);                                                   \
          } else { /* A^T|^H B and B^T|^H A */                                                                                \
            const T ai = a[p + i * lda], aj = a[p + j * lda], bi = b[p + i * ldb], bj = b[p + j * ldb];                       \
            s1 += (her ? CJ(ai) : ai) * bj; s2 += (her ? CJ(bi) : bi) * aj;                                                   \
          }                                                                                                                   \
        }                                                                                                                     \
        c[i + j * ldc] += alpha * s1 + alpha2 * s2;                                                                           \
      }                                                                                                                       \
    return her ? 1 : 0;                                                                                                       \
  }

L3_IMPL(s, float, float, 0, CJ_float, "S")
L3_IMPL(d, double, double, 0, CJ_double, "D")
L3_IMPL(c, cfloat, float, 1, CJ_cfloat, "C")
L3_IMPL(z, cdouble, double, 1, CJ_cdouble, "Z")

#define ABC_ARGS(R) const char* c1, const char* c2, const int* d1, const int* d2, const R* alpha, const R* a, const int* lda, const R* b, const int* ldb, const R* beta, R* c, const int* ldc
#define ABC_PASS c1, c2, d1, d2, alpha, a, lda, b, ldb, beta, c, ldc
int oracle_ssymm_(ABC_ARGS(float)) { return symm_s(0, "SSYMM ", ABC_PASS); }
int oracle_dsymm_(ABC_ARGS(double)) { return symm_d(0, "DSYMM ", ABC_PASS); }
int oracle_csymm_(ABC_ARGS(float)) { return symm_c(0, "CSYMM ", ABC_PASS); }
int oracle_zsymm_(ABC_ARGS(double)) { return symm_z(0, "ZSYMM ", ABC_PASS); }
int oracle_chemm_(ABC_ARGS(float)) { return symm_c(1, "CHEMM ", ABC_PASS); }
int oracle_zhemm_(ABC_ARGS(double)) { return symm_z(1, "ZHEMM ", ABC_PASS); }
int oracle_ssyr2k_(ABC_ARGS(float)) { return r2k_s(0, "SSYR2K", ABC_PASS); }
int oracle_dsyr2k_(ABC_ARGS(double)) { return r2k_d(0, "DSYR2K", ABC_PASS); }
int oracle_csyr2k_(ABC_ARGS(float)) { return r2k_c(0, "CSYR2K", ABC_PASS); }
int oracle_zsyr2k_(ABC_ARGS(double)) { return r2k_z(0, "ZSYR2K", ABC_PASS); }
int oracle_cher2k_(ABC_ARGS(float)) { return r2k_c(1, "CHER2K", ABC_PASS); }
int oracle_zher2k_(ABC_ARGS(double)) { return r2k_z(1, "ZHER2K", ABC_PASS); }
