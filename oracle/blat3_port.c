/* oracle/blat3_port.c -- CPU ORACLE, TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * C restatement of the GEMM part of the netlib level-3 BLAS testers that the reference ships and runs against
 * its blas/ library (no Fortran compiler exists in this image, so blas/testing/{s,d,c,z}blat3.f cannot be built):
 *   xBEG   generators            dblat3.f:2723-2768, zblat3.f:3344-3395 (LCG: i = i*891 mod 1000, every 5th skipped)
 *   xMAKE  'GE' matrices         dblat3.f:2395-2507 (zero column n/2 when n > 3, ROGUE = -1e10 in the ld padding)
 *   xCHK1  the GEMM sweep        dblat3.f:395-675   (dims 0 1 2 3 5 9, NMAX = 65, ld = dim+1, N/T/C x N/T/C,
 *                                                    alpha/beta from {s,d,c,z}blat3.dat:10-14)
 *   xMMCH  the numeric check     dblat3.f:2508-2627 (ratio = |ct-cc| / (eps*G) < THRESH = 16; fatal if
 *                                                    ratio*sqrt(eps) >= 1), working precision = the scalar's
 *   LxE / LxERES                 dblat3.f:2628-2722 (every input unchanged, padding rows of C unchanged)
 *   xCHKE  GEMM error exits      dblat3.f:1889-1972 + XERBLA/CHKXER :2771-2850
 * The routine under test is passed in as a function pointer, so the same checker runs against the oracle port,
 * the reference library (oracle/_ref) and the sm_100a library.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

#define NMAX 65
static const int k_idim[6] = {0, 1, 2, 3, 5, 9};
static const char k_ich[3] = {'N', 'T', 'C'};

/* ---- xBEG -------------------------------------------------------------------------------------------- */
static int g_i = 7, g_j = 7, g_ic = 0;
void oracle_blat3_reset(void) { g_i = 7; g_j = 7; g_ic = 0; }
static void beg_step(void) {
  g_ic += 1;
  for (;;) {
    g_i = g_i * 891; g_j = g_j * 457;
    g_i -= 1000 * (g_i / 1000); g_j -= 1000 * (g_j / 1000);
    if (g_ic >= 5) { g_ic = 0; continue; }
    break;
  }
}
double oracle_blat3_dbeg(void) { beg_step(); return (g_i - 500) / 1001.0; }
void oracle_blat3_zbeg(double* re, double* im) { beg_step(); *re = (g_i - 500) / 1001.0; *im = (g_j - 500) / 1001.0; }

/* ---- XERBLA replacement of the tester (dblat3.f:2818-2850): records instead of printing when armed --------- */
static int g_armed = 0, g_infot = 0, g_lerr = 0, g_ok = 1;
static char g_srnamt[8] = "";
static int tester_xerbla(const char* name, int* info, int len) {
  (void)len;
  if (!g_armed) {
    printf("Eigen BLAS ERROR #%i: %s\n", *info, name);
    return 0;
  }
  g_lerr = 1;
  if (*info != g_infot) g_ok = 0;
  if (strncmp(name, g_srnamt, 6) != 0) g_ok = 0;
  return 0;
}
/* exported so that a library resolving xerbla_ by symbol (the reference's weak definition, blas/xerbla.cpp:15,
 * and the sm_100a library's) binds to the tester's version when liboracle.so is loaded RTLD_GLOBAL first */
int xerbla_(const char* name, int* info, int len) { return tester_xerbla(name, info, len); }

/* arm / disarm the tester's XERBLA from a test written in another language (used for the ?syrk_/?herk_ error exits,
 * dblat3.f:2232-2296 DSYRK block of DCHKE): expect(name, infot) arms, result() returns 1 iff XERBLA was entered
 * with that name and info since, and disarms. */
void oracle_xerbla_expect(const char* name6, int infot) {
  memcpy(g_srnamt, name6, 6); g_srnamt[6] = 0;
  g_infot = infot; g_lerr = 0; g_ok = 1; g_armed = 1;
}
int oracle_xerbla_result(void) {
  const int r = g_lerr && g_ok;
  g_armed = 0;
  return r;
}

/* ---- storage helpers: every matrix is kept as (re,im) double pairs; typed copies are made for the call ------ */
typedef struct { int type, cplx, dbl; double eps; size_t esz; } tinfo;
static tinfo make_tinfo(int type) {
  tinfo t;
  t.type = type; t.cplx = type >= 2; t.dbl = (type == ORACLE_D || type == ORACLE_Z);
  t.eps = t.dbl ? 2.220446049250313e-16 : 1.1920928955078125e-07; /* EPSILON(ZERO) */
  t.esz = (t.dbl ? 8 : 4) * (t.cplx ? 2 : 1);
  return t;
}
static void put(const tinfo* t, void* buf, long idx, double re, double im) {
  if (t->dbl) { if (t->cplx) { ((double*)buf)[2 * idx] = re; ((double*)buf)[2 * idx + 1] = im; } else ((double*)buf)[idx] = re; }
  else { if (t->cplx) { ((float*)buf)[2 * idx] = (float)re; ((float*)buf)[2 * idx + 1] = (float)im; } else ((float*)buf)[idx] = (float)re; }
}
static void get(const tinfo* t, const void* buf, long idx, double* re, double* im) {
  if (t->dbl) { if (t->cplx) { *re = ((const double*)buf)[2 * idx]; *im = ((const double*)buf)[2 * idx + 1]; } else { *re = ((const double*)buf)[idx]; *im = 0; } }
  else { if (t->cplx) { *re = ((const float*)buf)[2 * idx]; *im = ((const float*)buf)[2 * idx + 1]; } else { *re = ((const float*)buf)[idx]; *im = 0; } }
}

/* xMAKE('GE'): A is the NMAX-strided master copy (values already rounded to the working precision), AA the
 * lda-strided array handed to the routine, padded with ROGUE */
static void make_ge(const tinfo* t, int m, int n, double* A /* [NMAX*NMAX*2] */, void* AA, int lda) {
  for (int j = 1; j <= n; ++j)
    for (int i = 1; i <= m; ++i) {
      double re, im = 0;
      if (t->cplx) oracle_blat3_zbeg(&re, &im); else re = oracle_blat3_dbeg();
      if (!t->dbl) { /* SBEG/CBEG divide in single precision */
        re = (double)((float)(g_i - 500) / 1001.0f);
        if (t->cplx) im = (double)((float)(g_j - 500) / 1001.0f);
      }
      if (i != j && n > 3 && j == n / 2) { re = 0; im = 0; }
      A[2 * ((i - 1) + (j - 1) * NMAX)] = re;
      A[2 * ((i - 1) + (j - 1) * NMAX) + 1] = im;
    }
  for (int j = 1; j <= n; ++j) {
    for (int i = 1; i <= m; ++i)
      put(t, AA, (i - 1) + (long)(j - 1) * lda, A[2 * ((i - 1) + (j - 1) * NMAX)], A[2 * ((i - 1) + (j - 1) * NMAX) + 1]);
    for (int i = m + 1; i <= lda; ++i) put(t, AA, (i - 1) + (long)(j - 1) * lda, -1.0e10, t->cplx ? 1.0e10 : 0.0);
  }
}

/* xMMCH in the working precision of the type (float arithmetic for s/c, double for d/z) */
#define MMCH_IMPL(NAME, W)                                                                                      \
  static double NAME(const tinfo* t, char transa, char transb, int m, int n, int kk, const double* alpha,       \
                     const double* A, const double* B, const double* beta, const double* C, const void* CC,     \
                     int ldcc, int* fatal) {                                                                    \
    const int trana = transa != 'N', tranb = transb != 'N';                                                     \
    const int ctrana = transa == 'C' && t->cplx, ctranb = transb == 'C' && t->cplx;                             \
    const W eps = (W)t->eps;                                                                                    \
    W err = 0;                                                                                                  \
    for (int j = 0; j < n; ++j) {                                                                               \
      W ctr[NMAX], cti[NMAX], g[NMAX];                                                                          \
      for (int i = 0; i < m; ++i) { ctr[i] = 0; cti[i] = 0; g[i] = 0; }                                          \
      for (int k = 0; k < kk; ++k)                                                                              \
        for (int i = 0; i < m; ++i) {                                                                           \
          const double* pa = trana ? &A[2 * (k + i * NMAX)] : &A[2 * (i + k * NMAX)];                           \
          const double* pb = tranb ? &B[2 * (j + k * NMAX)] : &B[2 * (k + j * NMAX)];                           \
          const W ar = (W)pa[0], ai = ctrana ? -(W)pa[1] : (W)pa[1];                                            \
          const W br = (W)pb[0], bi = ctranb ? -(W)pb[1] : (W)pb[1];                                            \
          ctr[i] = ctr[i] + (ar * br - ai * bi);                                                                \
          cti[i] = cti[i] + (ar * bi + ai * br);                                                                \
          g[i] = g[i] + (W)(fabs((double)ar) + fabs((double)ai)) * (W)(fabs((double)br) + fabs((double)bi));    \
        }                                                                                                       \
      for (int i = 0; i < m; ++i) {                                                                             \
        const W cr = (W)C[2 * (i + j * NMAX)], ci = (W)C[2 * (i + j * NMAX) + 1];                               \
        const W alr = (W)alpha[0], ali = (W)alpha[1], ber = (W)beta[0], bei = (W)beta[1];                       \
        const W nr_ = (alr * ctr[i] - ali * cti[i]) + (ber * cr - bei * ci);                                    \
        const W ni_ = (alr * cti[i] + ali * ctr[i]) + (ber * ci + bei * cr);                                    \
        ctr[i] = nr_; cti[i] = ni_;                                                                             \
        g[i] = (W)(fabs((double)alr) + fabs((double)ali)) * g[i] +                                              \
               (W)(fabs((double)ber) + fabs((double)bei)) * (W)(fabs((double)cr) + fabs((double)ci));           \
      }                                                                                                         \
      for (int i = 0; i < m; ++i) {                                                                             \
        double ccr, cci;                                                                                        \
        get(t, CC, i + (long)j * ldcc, &ccr, &cci);                                                             \
        W erri = (W)(fabs((double)(ctr[i] - (W)ccr)) + fabs((double)(cti[i] - (W)cci))) / eps;                  \
        if (g[i] != 0) erri = erri / g[i];                                                                      \
        if (!(erri <= err)) err = erri; /* also propagates NaN */                                               \
        if (!((double)err * sqrt((double)eps) < 1.0)) { *fatal = 1; return (double)err; }                       \
      }                                                                                                         \
    }                                                                                                           \
    return (double)err;                                                                                         \
  }
MMCH_IMPL(mmch_f, float)
MMCH_IMPL(mmch_d, double)

static void values(const tinfo* t, double alf[3][2], double bet[3][2]) {
  /* {s,d}blat3.dat:11,13 and {c,z}blat3.dat:12,14 */
  const double ar[3] = {0.0, 1.0, 0.7}, br[3] = {0.0, 1.0, 1.3};
  for (int i = 0; i < 3; ++i) { alf[i][0] = ar[i]; alf[i][1] = 0; bet[i][0] = br[i]; bet[i][1] = 0; }
  if (t->cplx) { alf[2][1] = -0.9; bet[2][1] = -1.1; }
  if (!t->dbl) for (int i = 0; i < 3; ++i) for (int c = 0; c < 2; ++c) { alf[i][c] = (double)(float)alf[i][c]; bet[i][c] = (double)(float)bet[i][c]; }
}

void oracle_blat3_chk1(int type, oracle_gemm_fn gemm, oracle_blat3_report* rep) {
  const tinfo t = make_tinfo(type);
  double* A = (double*)calloc(2 * NMAX * NMAX, sizeof(double));
  double* B = (double*)calloc(2 * NMAX * NMAX, sizeof(double));
  double* C = (double*)calloc(2 * NMAX * NMAX, sizeof(double));
  void *AA = malloc(t.esz * NMAX * NMAX), *AS = malloc(t.esz * NMAX * NMAX);
  void *BB = malloc(t.esz * NMAX * NMAX), *BS = malloc(t.esz * NMAX * NMAX);
  void *CC = malloc(t.esz * NMAX * NMAX), *CS = malloc(t.esz * NMAX * NMAX);
  double alf[3][2], bet[3][2];
  values(&t, alf, bet);
  memset(rep, 0, sizeof(*rep));
  oracle_blat3_reset();
  g_armed = 1; g_ok = 1; g_lerr = 0; g_infot = 0; /* a valid call must not reach XERBLA ("OK" in COMMON /INFOC/) */
  for (int im = 0; im < 6 && !rep->fatal; ++im) {
    const int m = k_idim[im];
    for (int in = 0; in < 6 && !rep->fatal; ++in) {
      const int n = k_idim[in];
      int ldc = m; if (ldc < NMAX) ldc += 1;
      if (ldc > NMAX) continue;
      const long lcc = (long)ldc * n;
      const int null = n <= 0 || m <= 0;
      for (int ik = 0; ik < 6 && !rep->fatal; ++ik) {
        const int k = k_idim[ik];
        for (int ica = 0; ica < 3 && !rep->fatal; ++ica) {
          const char transa = k_ich[ica];
          const int trana = transa != 'N';
          const int ma = trana ? k : m, na = trana ? m : k;
          int lda = ma; if (lda < NMAX) lda += 1;
          if (lda > NMAX) continue;
          const long laa = (long)lda * na;
          make_ge(&t, ma, na, A, AA, lda);
          for (int icb = 0; icb < 3 && !rep->fatal; ++icb) {
            const char transb = k_ich[icb];
            const int tranb = transb != 'N';
            const int mb = tranb ? n : k, nb = tranb ? k : n;
            int ldb = mb; if (ldb < NMAX) ldb += 1;
            if (ldb > NMAX) continue;
            const long lbb = (long)ldb * nb;
            make_ge(&t, mb, nb, B, BB, ldb);
            for (int ia = 0; ia < 3 && !rep->fatal; ++ia)
              for (int ib = 0; ib < 3 && !rep->fatal; ++ib) {
                make_ge(&t, m, n, C, CC, ldc);
                rep->ncalls += 1;
                /* save every datum before the call */
                char tas = transa, tbs = transb;
                int ms = m, ns = n, ks = k, ldas = lda, ldbs = ldb, ldcs = ldc;
                unsigned char als[16], bls[16], alv[16], bev[16];
                put(&t, alv, 0, alf[ia][0], alf[ia][1]); put(&t, bev, 0, bet[ib][0], bet[ib][1]);
                memcpy(als, alv, t.esz); memcpy(bls, bev, t.esz);
                memcpy(AS, AA, t.esz * laa); memcpy(BS, BB, t.esz * lbb); memcpy(CS, CC, t.esz * lcc);
                char ta = transa, tb = transb;
                int mm = m, nn = n, kk = k, la = lda, lb = ldb, lc = ldc;
                gemm(&ta, &tb, &mm, &nn, &kk, alv, AA, &la, BB, &lb, bev, CC, &lc);
                if (!g_ok || g_lerr) {
                  rep->fatal = 1; snprintf(rep->msg, sizeof rep->msg, "ERROR-EXIT TAKEN ON VALID CALL");
                  break;
                }
                int isame[14]; memset(isame, 0, sizeof isame);
                isame[1] = ta == tas; isame[2] = tb == tbs; isame[3] = ms == mm; isame[4] = ns == nn;
                isame[5] = ks == kk; isame[6] = memcmp(als, alv, t.esz) == 0;
                isame[7] = memcmp(AS, AA, t.esz * laa) == 0; isame[8] = ldas == la;
                isame[9] = memcmp(BS, BB, t.esz * lbb) == 0; isame[10] = ldbs == lb;
                isame[11] = memcmp(bls, bev, t.esz) == 0;
                if (null) isame[12] = memcmp(CS, CC, t.esz * lcc) == 0;
                else { /* LxERES('GE'): rows m+1..ldc of every column untouched */
                  isame[12] = 1;
                  for (int j = 0; j < n; ++j)
                    if (memcmp((char*)CS + t.esz * (m + (long)j * ldc), (char*)CC + t.esz * (m + (long)j * ldc), t.esz * (ldc - m)) != 0) isame[12] = 0;
                }
                isame[13] = ldcs == lc;
                for (int i = 1; i <= 13; ++i)
                  if (!isame[i] && !rep->fatal) {
                    rep->fatal = 1; rep->bad_param = i;
                    snprintf(rep->msg, sizeof rep->msg, "PARAMETER NUMBER %d WAS CHANGED INCORRECTLY", i);
                  }
                if (!rep->fatal && !null) {
                  int fatal = 0;
                  const double err = t.dbl ? mmch_d(&t, transa, transb, m, n, k, alf[ia], A, B, bet[ib], C, CC, ldc, &fatal)
                                           : mmch_f(&t, transa, transb, m, n, k, alf[ia], A, B, bet[ib], C, CC, ldc, &fatal);
                  if (err > rep->errmax || err != err) rep->errmax = err;
                  if (fatal) { rep->fatal = 1; snprintf(rep->msg, sizeof rep->msg, "COMPUTED RESULT IS LESS THAN HALF ACCURATE"); }
                }
                if (rep->fatal) {
                  const size_t l = strlen(rep->msg);
                  snprintf(rep->msg + l, sizeof rep->msg - l, " [call %d: ('%c','%c',%d,%d,%d, alpha#%d, A,%d, B,%d, beta#%d, C,%d)]",
                           rep->ncalls, transa, transb, m, n, k, ia, lda, ldb, ib, ldc);
                }
              }
          }
        }
      }
    }
  }
  g_armed = 0;
  free(A); free(B); free(C); free(AA); free(AS); free(BB); free(BS); free(CC); free(CS);
}

/* xCHKE, GEMM block (dblat3.f:1889-1972): 28 illegal calls; CHKXER requires XERBLA to have been entered with the
 * expected INFO and routine name. */
int oracle_blat3_chke(int type, oracle_gemm_fn gemm, oracle_install_xerbla_fn install, char* log, int loglen) {
  static const struct { int infot; char ta, tb; int m, n, k, lda, ldb, ldc; } cases[28] = {
      {1, '/', 'N', 0, 0, 0, 1, 1, 1},  {1, '/', 'T', 0, 0, 0, 1, 1, 1},  {2, 'N', '/', 0, 0, 0, 1, 1, 1},
      {2, 'T', '/', 0, 0, 0, 1, 1, 1},  {3, 'N', 'N', -1, 0, 0, 1, 1, 1}, {3, 'N', 'T', -1, 0, 0, 1, 1, 1},
      {3, 'T', 'N', -1, 0, 0, 1, 1, 1}, {3, 'T', 'T', -1, 0, 0, 1, 1, 1}, {4, 'N', 'N', 0, -1, 0, 1, 1, 1},
      {4, 'N', 'T', 0, -1, 0, 1, 1, 1}, {4, 'T', 'N', 0, -1, 0, 1, 1, 1}, {4, 'T', 'T', 0, -1, 0, 1, 1, 1},
      {5, 'N', 'N', 0, 0, -1, 1, 1, 1}, {5, 'N', 'T', 0, 0, -1, 1, 1, 1}, {5, 'T', 'N', 0, 0, -1, 1, 1, 1},
      {5, 'T', 'T', 0, 0, -1, 1, 1, 1}, {8, 'N', 'N', 2, 0, 0, 1, 1, 2},  {8, 'N', 'T', 2, 0, 0, 1, 1, 2},
      {8, 'T', 'N', 0, 0, 2, 1, 2, 1},  {8, 'T', 'T', 0, 0, 2, 1, 1, 1},  {10, 'N', 'N', 0, 0, 2, 1, 1, 1},
      {10, 'T', 'N', 0, 0, 2, 2, 1, 1}, {10, 'N', 'T', 0, 2, 0, 1, 1, 1}, {10, 'T', 'T', 0, 2, 0, 1, 1, 1},
      {13, 'N', 'N', 2, 0, 0, 2, 1, 1}, {13, 'N', 'T', 2, 0, 0, 2, 1, 1}, {13, 'T', 'N', 2, 0, 0, 1, 1, 1},
      {13, 'T', 'T', 2, 0, 0, 1, 1, 1}};
  static const char* names[4] = {"SGEMM ", "DGEMM ", "CGEMM ", "ZGEMM "};
  const tinfo t = make_tinfo(type);
  unsigned char A[16 * 4], B[16 * 4], C[16 * 4], alpha[16], beta[16];
  memset(A, 0, sizeof A); memset(B, 0, sizeof B); memset(C, 0, sizeof C);
  put(&t, alpha, 0, 1.0, 0.0); put(&t, beta, 0, 2.0, 0.0);
  if (install) install(tester_xerbla);
  int failed = 0, pos = 0;
  if (log && loglen > 0) log[0] = 0;
  memcpy(g_srnamt, names[type], 7);
  g_armed = 1;
  for (int i = 0; i < 28; ++i) {
    g_infot = cases[i].infot; g_lerr = 0; g_ok = 1;
    char ta = cases[i].ta, tb = cases[i].tb;
    int m = cases[i].m, n = cases[i].n, k = cases[i].k, lda = cases[i].lda, ldb = cases[i].ldb, ldc = cases[i].ldc;
    unsigned char C0[sizeof C]; memcpy(C0, C, sizeof C);
    gemm(&ta, &tb, &m, &n, &k, alpha, A, &lda, B, &ldb, beta, C, &ldc);
    const int bad = !g_lerr || !g_ok || memcmp(C0, C, sizeof C) != 0;
    if (bad) {
      failed += 1;
      if (log && pos < loglen - 80)
        pos += snprintf(log + pos, loglen - pos, "case %d: ILLEGAL VALUE OF PARAMETER NUMBER %d %s\n", i, cases[i].infot,
                        !g_lerr ? "NOT DETECTED" : "DETECTED WITH WRONG INFO/NAME OR C MODIFIED");
    }
  }
  g_armed = 0;
  if (install) install(NULL);
  return failed;
}
