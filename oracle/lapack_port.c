/* oracle/lapack_port.c -- CPU ORACLE, TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Plain-C (C99 complex) restatement of the reference's blocked factorizations:
 *   EIGEN_LAPACK_FUNC(potrf)   lapack/cholesky.cpp:14-38   -> llt_inplace<Scalar,UpLo>::blocked / unblocked
 *                                                             (Eigen/src/Cholesky/LLT.h:299-360; Upper = Lower on the
 *                                                             TRANSPOSED view, no conjugation, LLT.h:363-380)
 *   EIGEN_LAPACK_FUNC(getrf)   lapack/lu.cpp:14-42         -> partial_lu_impl::blocked_lu / unblocked_lu
 *                                                             (Eigen/src/LU/PartialPivLU.h:361-496)
 * Same block sizes (size/8 rounded down to 16, clamped to [8,128] for LLT, [8,256] and 16 inside a panel for LU), same
 * pivot rule (largest |a|, first index on ties; an exactly-zero column is recorded and skipped), same info values.
 * The inner sums are plain sequential chains, so values agree with the compiled reference to a few eps, not bitwise
 * (tests/test_oracle_pin_lapack.py pins that, together with identical pivots and info).
 */
#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include "oracle.h"

int oracle_call_xerbla_public(const char* name, int* info); /* gebp_port.c */
static long lmin(long a, long b) { return a < b ? a : b; }
static long lmax(long a, long b) { return a > b ? a : b; }
typedef float _Complex cfloat;
typedef double _Complex cdouble;
#define ABS_float(x) fabsf(x)
#define ABS_double(x) fabs(x)
#define ABS_cfloat(x) cabsf(x)
#define ABS_cdouble(x) cabs(x)
#define CJ_float(x) (x)
#define CJ_double(x) (x)
#define CJ_cfloat(x) conjf(x)
#define CJ_cdouble(x) conj(x)

/* EL(i,j): element (i,j) of the lower-canonical view (the matrix itself for Lower, its transpose for Upper) */
#define LAPACK_IMPL(SFX, T, R, TN, NM)                                                                                        \
  static long llt_unblocked_##SFX(T* a, long lda, long n, int upper) { /* LLT.h:299-323 */                                    \
    _Pragma("GCC diagnostic ignored \"-Wunused-value\"")                                                                      \
    for (long k = 0; k < n; ++k) {                                                                                            \
      T* akk = upper ? a + k + k * lda : a + k + k * lda;                                                                     \
      R x = (R)creal((cdouble)*akk);                                                                                          \
      for (long j = 0; j < k; ++j) { const T v = upper ? a[j + k * lda] : a[k + j * lda]; x -= (R)creal((cdouble)(v * CJ_##TN(v))); } \
      if (x <= (R)0) return k;                                                                                                \
      x = (R)sqrt((double)x);                                                                                                 \
      *akk = (T)x;                                                                                                            \
      for (long i = k + 1; i < n; ++i) {                                                                                      \
        T* aik = upper ? a + k + i * lda : a + i + k * lda;                                                                   \
        T acc = *aik;                                                                                                         \
        for (long j = 0; j < k; ++j) {                                                                                        \
          const T aij = upper ? a[j + i * lda] : a[i + j * lda], akj = upper ? a[j + k * lda] : a[k + j * lda];              \
          acc -= aij * CJ_##TN(akj);                                                                                          \
        }                                                                                                                     \
        *aik = acc / (T)x;                                                                                                    \
      }                                                                                                                       \
    }                                                                                                                         \
    return -1;                                                                                                                \
  }                                                                                                                           \
  static long llt_blocked_##SFX(T* a, long lda, long size, int upper) { /* LLT.h:325-360 */                                   \
    if (size < 32) return llt_unblocked_##SFX(a, lda, size, upper);                                                           \
    long bsz = size / 8;                                                                                                      \
    bsz = (bsz / 16) * 16;                                                                                                    \
    bsz = lmin(lmax(bsz, 8), 128);                                                                                            \
    for (long k = 0; k < size; k += bsz) {                                                                                    \
      const long bs = lmin(bsz, size - k), rs = size - k - bs;                                                                \
      const long ret = llt_unblocked_##SFX(a + k + k * lda, lda, bs, upper);                                                  \
      if (ret >= 0) return k + ret;                                                                                           \
      /* canonical (lower) view: A21 = A21 * L11^-H, then A22 -= A21 * A21^H on the lower triangle */                         \
      for (long i = 0; i < rs; ++i)                                                                                           \
        for (long j = 0; j < bs; ++j) {                                                                                       \
          T* x = upper ? a + (k + j) + (k + bs + i) * lda : a + (k + bs + i) + (k + j) * lda;                                 \
          T acc = *x;                                                                                                         \
          for (long p = 0; p < j; ++p) {                                                                                      \
            const T xp = upper ? a[(k + p) + (k + bs + i) * lda] : a[(k + bs + i) + (k + p) * lda];                           \
            const T ljp = upper ? a[(k + p) + (k + j) * lda] : a[(k + j) + (k + p) * lda];                                    \
            acc -= xp * CJ_##TN(ljp);                                                                                         \
          }                                                                                                                   \
          const T ljj = a[(k + j) + (k + j) * lda];                                                                           \
          *x = acc / CJ_##TN(ljj);                                                                                            \
        }                                                                                                                     \
      for (long j = 0; j < rs; ++j)                                                                                           \
        for (long i = j; i < rs; ++i) {                                                                                       \
          T acc = (T)0;                                                                                                       \
          for (long p = 0; p < bs; ++p) {                                                                                     \
            const T aip = upper ? a[(k + p) + (k + bs + i) * lda] : a[(k + bs + i) + (k + p) * lda];                          \
            const T ajp = upper ? a[(k + p) + (k + bs + j) * lda] : a[(k + bs + j) + (k + p) * lda];                          \
            acc += aip * CJ_##TN(ajp);                                                                                        \
          }                                                                                                                   \
          T* c = upper ? a + (k + bs + j) + (k + bs + i) * lda : a + (k + bs + i) + (k + bs + j) * lda;                       \
          *c -= acc;                                                                                                          \
        }                                                                                                                     \
    }                                                                                                                         \
    return -1;                                                                                                                \
  }                                                                                                                           \
  int oracle_##SFX##potrf_(const char* uplo, const int* n, R* pa, const int* lda, int* info) {                                \
    const int up = (*uplo == 'U' || *uplo == 'u') ? 1 : (*uplo == 'L' || *uplo == 'l') ? 0 : -1;                              \
    *info = 0;                                                                                                                \
    if (up < 0) *info = -1;                                                                                                   \
    else if (*n < 0) *info = -2;                                                                                              \
    else if (*lda < (*n > 1 ? *n : 1)) *info = -4;                                                                            \
    if (*info != 0) { int e = -*info; return oracle_call_xerbla_public(NM "POTRF", &e); }                                     \
    const long ret = llt_blocked_##SFX((T*)pa, *lda, *n, up);                                                                 \
    if (ret >= 0) *info = (int)ret + 1;                                                                                       \
    return 0;                                                                                                                 \
  }                                                                                                                           \
  static long lu_unblocked_##SFX(T* lu, long ld, long rows, long cols, int* piv) { /* PartialPivLU.h:361-408 */               \
    const long size = lmin(rows, cols);                                                                                       \
    long first_zero = -1;                                                                                                     \
    for (long k = 0; k < size; ++k) {                                                                                         \
      long best = k;                                                                                                          \
      R bscore = ABS_##TN(lu[k + k * ld]);                                                                                    \
      for (long i = k + 1; i < rows; ++i) { const R sc = ABS_##TN(lu[i + k * ld]); if (sc > bscore) { bscore = sc; best = i; } } \
      piv[k] = (int)best;                                                                                                     \
      if (bscore != (R)0) {                                                                                                   \
        if (best != k) for (long j = 0; j < cols; ++j) { const T t = lu[k + j * ld]; lu[k + j * ld] = lu[best + j * ld]; lu[best + j * ld] = t; } \
        const T d = lu[k + k * ld];                                                                                           \
        for (long i = k + 1; i < rows; ++i) lu[i + k * ld] /= d;                                                              \
      } else if (first_zero == -1) first_zero = k;                                                                            \
      if (k < rows - 1)                                                                                                       \
        for (long j = k + 1; j < cols; ++j) { const T u = lu[k + j * ld]; for (long i = k + 1; i < rows; ++i) lu[i + j * ld] -= lu[i + k * ld] * u; } \
    }                                                                                                                         \
    return first_zero;                                                                                                        \
  }                                                                                                                           \
  static long lu_blocked_##SFX(T* lu, long ld, long rows, long cols, int* piv, long max_bs) { /* PartialPivLU.h:424-496 */    \
    const long size = lmin(rows, cols);                                                                                       \
    if (size <= 16) return lu_unblocked_##SFX(lu, ld, rows, cols, piv);                                                       \
    long bsz = size / 8;                                                                                                      \
    bsz = (bsz / 16) * 16;                                                                                                    \
    bsz = lmin(lmax(bsz, 8), max_bs);                                                                                         \
    long first_zero = -1;                                                                                                     \
    for (long k = 0; k < size; k += bsz) {                                                                                    \
      const long bs = lmin(size - k, bsz), trows = rows - k - bs, tsize = size - k - bs;                                      \
      const long ret = lu_blocked_##SFX(lu + k + k * ld, ld, trows + bs, bs, piv + k, 16);                                    \
      if (ret >= 0 && first_zero == -1) first_zero = k + ret;                                                                 \
      for (long i = k; i < k + bs; ++i) { /* globalise the pivots, apply them to A_0 (columns < k) */                         \
        const long p = (piv[i] += (int)k);                                                                                    \
        if (p != i) for (long j = 0; j < k; ++j) { const T t = lu[i + j * ld]; lu[i + j * ld] = lu[p + j * ld]; lu[p + j * ld] = t; } \
      }                                                                                                                       \
      if (trows) {                                                                                                            \
        for (long i = k; i < k + bs; ++i) { /* A_2: columns [k+bs, k+bs+tsize) */                                             \
          const long p = piv[i];                                                                                              \
          if (p != i) for (long j = k + bs; j < k + bs + tsize; ++j) { const T t = lu[i + j * ld]; lu[i + j * ld] = lu[p + j * ld]; lu[p + j * ld] = t; } \
        }                                                                                                                     \
        for (long j = k + bs; j < k + bs + tsize; ++j) { /* A12 = L11^-1 A12 (unit lower) */                                  \
          for (long i = k; i < k + bs; ++i) {                                                                                 \
            T acc = lu[i + j * ld];                                                                                           \
            for (long p = k; p < i; ++p) acc -= lu[i + p * ld] * lu[p + j * ld];                                              \
            lu[i + j * ld] = acc;                                                                                             \
          }                                                                                                                   \
          for (long i = k + bs; i < rows; ++i) { /* A22 -= A21 * A12 */                                                       \
            T acc = (T)0;                                                                                                     \
            for (long p = k; p < k + bs; ++p) acc += lu[i + p * ld] * lu[p + j * ld];                                         \
            lu[i + j * ld] -= acc;                                                                                            \
          }                                                                                                                   \
        }                                                                                                                     \
      }                                                                                                                       \
    }                                                                                                                         \
    return first_zero;                                                                                                        \
  }                                                                                                                           \
  int oracle_##SFX##getrf_(const int* m, const int* n, R* pa, const int* lda, int* ipiv, int* info) {                         \
    *info = 0;                                                                                                                \
    if (*m < 0) *info = -1;                                                                                                   \
    else if (*n < 0) *info = -2;                                                                                              \
    else if (*lda < (*m > 1 ? *m : 1)) *info = -4;                                                                            \
    if (*info != 0) { int e = -*info; return oracle_call_xerbla_public(NM "GETRF", &e); }                                     \
    if (*m == 0 || *n == 0) return 0;                                                                                         \
    const long ret = lu_blocked_##SFX((T*)pa, *lda, *m, *n, ipiv, 256);                                                       \
    for (long i = 0; i < lmin(*m, *n); ++i) ipiv[i]++;                                                                        \
    if (ret >= 0) *info = (int)ret + 1;                                                                                       \
    return 0;                                                                                                                 \
  }

LAPACK_IMPL(s, float, float, float, "S")
LAPACK_IMPL(d, double, double, double, "D")
LAPACK_IMPL(c, cfloat, float, cfloat, "C")
LAPACK_IMPL(z, cdouble, double, cdouble, "Z")
