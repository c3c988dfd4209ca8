/* oracle/hp_ref.c -- CPU ORACLE, TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * High-precision GEMM reference: every dot product is accumulated in x87 long double (64-bit mantissa) from
 * exact products of the inputs, column at a time like the netlib checker xMMCH
 * (blas/testing/dblat3.f:2508-2627, zblat3.f ZMMCH), which also yields the gauge
 *     G(i,j) = |alpha| * sum_k |a_ik||b_kj| + |beta| * |c_ij|          (ABS1 = |re|+|im| for complex)
 * against which xBLAT3 normalises errors.  BASELINE.json's tolerance (relative Frobenius error <= c*k*eps vs a
 * long-double reference) is evaluated against this function.  Rows can be sampled (row_idx) so that N = 16384
 * products are checked in seconds.
 */
#include <math.h>
#include <stdlib.h>
#include "oracle.h"

typedef long double ld;

static int op_of(char x) {
  return (x == 'N' || x == 'n') ? 0 : (x == 'T' || x == 't') ? 1 : 2;
}

#define LOADER(NAME, T)                                                                                     \
  static inline void NAME(const void* p, int cplx, long idx, ld* re, ld* im) {                              \
    const T* q = (const T*)p;                                                                               \
    if (cplx) { *re = q[2 * idx]; *im = q[2 * idx + 1]; } else { *re = q[idx]; *im = 0; }                   \
  }
LOADER(load_f, float)
LOADER(load_d, double)

void oracle_hp_gemm(int type, char ta, char tb, int m, int n, int k, const void* alpha, const void* a, int lda,
                    const void* b, int ldb, const void* beta, const void* c, int ldc, const int* row_idx, int nrows,
                    double* out, double* gauge) {
  const int cplx = type >= 2, dbl = (type == ORACLE_D || type == ORACLE_Z);
  const int oa = op_of(ta), ob = op_of(tb);
  if (!row_idx) nrows = m;
  ld al_re, al_im, be_re, be_im;
  if (dbl) { load_d(alpha, cplx, 0, &al_re, &al_im); load_d(beta, cplx, 0, &be_re, &be_im); }
  else { load_f(alpha, cplx, 0, &al_re, &al_im); load_f(beta, cplx, 0, &be_re, &be_im); }
  const ld abs_al = cplx ? fabsl(al_re) + fabsl(al_im) : fabsl(al_re);
  const ld abs_be = cplx ? fabsl(be_re) + fabsl(be_im) : fabsl(be_re);
  const int beta_zero = (be_re == 0 && be_im == 0);
#pragma omp parallel for collapse(2) schedule(static)
  for (int j = 0; j < n; ++j)
    for (int r = 0; r < nrows; ++r) {
      const int i = row_idx ? row_idx[r] : r;
      ld sre = 0, sim = 0, g = 0;
      for (int p = 0; p < k; ++p) {
        ld are, aim, bre, bim;
        const long ia = oa == 0 ? (long)i + (long)p * lda : (long)p + (long)i * lda;
        const long ib = ob == 0 ? (long)p + (long)j * ldb : (long)j + (long)p * ldb;
        if (dbl) { load_d(a, cplx, ia, &are, &aim); load_d(b, cplx, ib, &bre, &bim); }
        else { load_f(a, cplx, ia, &are, &aim); load_f(b, cplx, ib, &bre, &bim); }
        if (oa == 2) aim = -aim;
        if (ob == 2) bim = -bim;
        sre += are * bre - aim * bim;
        sim += are * bim + aim * bre;
        g += (fabsl(are) + fabsl(aim)) * (fabsl(bre) + fabsl(bim));
      }
      ld cre = 0, cim = 0;
      if (!beta_zero) { /* beta == 0: C is never read (blas/level3_impl.h:64) */
        const long ic = (long)i + (long)j * ldc;
        if (dbl) load_d(c, cplx, ic, &cre, &cim); else load_f(c, cplx, ic, &cre, &cim);
      }
      const ld ore = al_re * sre - al_im * sim + be_re * cre - be_im * cim;
      const ld oim = al_re * sim + al_im * sre + be_re * cim + be_im * cre;
      const ld gg = abs_al * g + abs_be * (fabsl(cre) + fabsl(cim));
      const long io = (long)r + (long)j * nrows;
      /* results are always returned in double (also for s/c) so that fp32 errors are measured cleanly */
      if (cplx) { ((double*)out)[2 * io] = (double)ore; ((double*)out)[2 * io + 1] = (double)oim; }
      else ((double*)out)[io] = (double)ore;
      if (gauge) ((double*)gauge)[io] = (double)gg;
    }
}
