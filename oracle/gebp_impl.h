/* oracle/gebp_impl.h -- type-generic body of the GEMM oracle; TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Included four times by gebp_port.c with
 *   R      real type (float / double)
 *   NC     components per scalar (1 = real, 2 = complex, interleaved re,im)
 *   SFX    s / d / c / z
 *   FMA    fmaf / fma
 * It restates, for an AVX2+FMA build of the reference (32-byte packets):
 *   gemm_pack_lhs / gemm_pack_rhs      Eigen/src/Core/products/GeneralBlockPanelKernel.h:1688-2105
 *   gebp_kernel                        GeneralBlockPanelKernel.h:858-1669
 *   general_matrix_matrix_product::run GeneralMatrixMatrix.h:59-199 (sequential branch :155-198)
 *   EIGEN_BLAS_FUNC(gemm)              blas/level3_impl.h:12-76
 */
#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name##_, SFX)

/* packet geometry: gebp_traits (GeneralBlockPanelKernel.h:369-380 real, :618-619 complex) */
#define PKT ((int)(ORACLE_SIMD_BYTES / (sizeof(R) * NC))) /* scalars per packet = LhsProgress */
#if NC == 1
#define MR (3 * PKT) /* FMA available => mr = 3 packets */
#else
#define MR (PKT)     /* complex x complex => mr = 1 packet */
#endif
#define NR 4

/* element (i,j) of a strided source: order 0 = ColMajor, 1 = RowMajor (const_blas_data_mapper, BlasUtil.h:158-268) */
static inline const R* FN(at)(const R* p, long stride, int order, long i, long j) {
  return p + NC * (order == 0 ? i + j * stride : j + i * stride);
}

/* gemm_pack_lhs (GeneralBlockPanelKernel.h:1688-1881): row groups of 3P, 2P, P, then single rows, each stored
 * k-major; ColMajor and RowMajor sources produce the same packed image.  conj folded in (:1707). */
static void FN(pack_lhs)(R* blockA, const R* lhs, long stride, int order, int conj, long depth, long rows) {
  const int P = PKT;
  long count = 0, i = 0;
  const long peeled3 = MR >= 3 * P ? (rows / (3 * P)) * (3 * P) : 0;
  const long peeled2 = MR >= 2 * P ? peeled3 + ((rows - peeled3) / (2 * P)) * (2 * P) : 0;
  const long peeled1 = MR >= 1 * P ? (rows / P) * P : 0;
  const long bounds[3] = {peeled3, peeled2, peeled1};
  for (int g = 0; g < 3; ++g) {
    const int w = (3 - g) * P;
    for (; i < bounds[g]; i += w)
      for (long k = 0; k < depth; ++k)
        for (int r = 0; r < w; ++r) {
          const R* s = FN(at)(lhs, stride, order, i + r, k);
          blockA[count++] = s[0];
#if NC == 2
          blockA[count++] = conj ? -s[1] : s[1];
#endif
        }
  }
  for (; i < rows; ++i)
    for (long k = 0; k < depth; ++k) {
      const R* s = FN(at)(lhs, stride, order, i, k);
      blockA[count++] = s[0];
#if NC == 2
      blockA[count++] = conj ? -s[1] : s[1];
#endif
    }
  (void)conj;
}

/* gemm_pack_rhs (GeneralBlockPanelKernel.h:1890-2105): groups of nr=4 columns stored k-major with the 4 values
 * of a k adjacent (:1958-1997), remaining columns one by one (:1999-2010). */
static void FN(pack_rhs)(R* blockB, const R* rhs, long stride, int order, int conj, long depth, long cols) {
  long count = 0;
  const long cols4 = (cols / 4) * 4;
  for (long j = 0; j < cols4; j += 4)
    for (long k = 0; k < depth; ++k)
      for (int c = 0; c < 4; ++c) {
        const R* s = FN(at)(rhs, stride, order, k, j + c);
        blockB[count++] = s[0];
#if NC == 2
        blockB[count++] = conj ? -s[1] : s[1];
#endif
      }
  for (long j = cols4; j < cols; ++j)
    for (long k = 0; k < depth; ++k) {
      const R* s = FN(at)(rhs, stride, order, k, j);
      blockB[count++] = s[0];
#if NC == 2
      blockB[count++] = conj ? -s[1] : s[1];
#endif
    }
  (void)conj;
}

/* packed-panel addressing used by gebp (blockA[i*strideA + ...], blockB[j2*strideB + ...], :935-940,1514-1530) */
static inline const R* FN(pa)(const R* blockA, long depth, long rows, long i, long* kstride) {
  /* find the group that holds row i */
  const int P = PKT;
  const long peeled3 = MR >= 3 * P ? (rows / (3 * P)) * (3 * P) : 0;
  const long peeled2 = MR >= 2 * P ? peeled3 + ((rows - peeled3) / (2 * P)) * (2 * P) : 0;
  const long peeled1 = (rows / P) * P;
  long g0; int w;
  if (i < peeled3) { w = 3 * P; g0 = (i / w) * w; }
  else if (i < peeled2) { w = 2 * P; g0 = peeled3 + ((i - peeled3) / w) * w; }
  else if (i < peeled1) { w = P; g0 = peeled2 + ((i - peeled2) / w) * w; }
  else { w = 1; g0 = i; }
  *kstride = (long)NC * w;
  return blockA + NC * (g0 * depth + (i - g0));
}
static inline const R* FN(pb)(const R* blockB, long depth, long cols, long j, long* kstride) {
  const long cols4 = (cols / 4) * 4;
  if (j < cols4) { const long g0 = (j / 4) * 4; *kstride = (long)NC * 4; return blockB + NC * (g0 * depth + (j - g0)); }
  *kstride = NC;
  return blockB + NC * (j * depth);
}

#if NC == 1
/* gebp_kernel, real scalars.  Rounding model of the packet paths 3Px4 / 2Px4 / 1Px4 and their single-column
 * tails (GeneralBlockPanelKernel.h:915-1512): one accumulator per C element, a strictly sequential FMA chain
 * over the kc block starting from 0 (traits.madd = pmadd, :435-448), then C = fma(acc, alpha, C) (traits.acc,
 * :450-453).  Rows beyond the last full packet take the "swapped" path (:1514-1668). */
static void FN(gebp)(R* res, long ldc, const R* blockA, const R* blockB, long rows, long depth, long cols, R alpha) {
  const int P = PKT;
  const long peeled1 = (rows / P) * P;
  const long cols4 = (cols / 4) * 4;
  for (long j = 0; j < cols; ++j)
    for (long i = 0; i < peeled1; ++i) {
      long as, bs;
      const R* a = FN(pa)(blockA, depth, rows, i, &as);
      const R* b = FN(pb)(blockB, depth, cols, j, &bs);
      R acc = 0;
      for (long k = 0; k < depth; ++k) acc = FMA(a[k * as], b[k * bs], acc);
      res[i + j * ldc] = FMA(acc, alpha, res[i + j * ldc]);
    }
  if (peeled1 == rows) return;
  /* tail rows x 4-column groups: SwappedTraits vector path (:1528-1616).  LhsProgress = P in {4, 8};
   * spk = max(1, P/4) depth steps per packet; four packet accumulators C0..C3 are cycled over the depth,
   * reduced as (C0+C1)+(C2+C3), the remainder continues on C0; P == 8 folds the two half packets
   * (predux_downto4) before an optional last odd depth step. */
  const int spk = P / 4 > 1 ? P / 4 : 1;
  const long endk = (depth / spk) * spk, endk4 = (depth / (spk * 4)) * (spk * 4);
  for (long j2 = 0; j2 < cols4; j2 += 4)
    for (long i = peeled1; i < rows; ++i)
      for (int c = 0; c < 4; ++c) {
        long as, bs;
        const R* a = FN(pa)(blockA, depth, rows, i, &as);
        const R* b = FN(pb)(blockB, depth, cols, j2 + c, &bs);
        R C[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
        long k = 0;
        for (; k < endk4; k += 4 * spk)
          for (int q = 0; q < 4; ++q)
            for (int g = 0; g < spk; ++g) {
              const long kk = k + q * spk + g;
              C[q][g] = FMA(b[kk * bs], a[kk * as], C[q][g]);
            }
        R C0[2];
        for (int g = 0; g < spk; ++g) C0[g] = (C[0][g] + C[1][g]) + (C[2][g] + C[3][g]);
        for (; k < endk; k += spk)
          for (int g = 0; g < spk; ++g)
            C0[g] = FMA(b[(k + g) * bs], a[(k + g) * as], C0[g]);
        R c0 = C0[0];
        if (spk == 2) {
          c0 = C0[0] + C0[1];
          if (depth - endk > 0)
            c0 = FMA(b[endk * bs], a[endk * as], c0);
        }
        res[i + (j2 + c) * ldc] = FMA(c0, alpha, res[i + (j2 + c) * ldc]);
      }
  /* tail rows x tail columns: scalar CJMADD chain, then res += alpha*C0 (:1650-1667) */
  for (long j = cols4; j < cols; ++j)
    for (long i = peeled1; i < rows; ++i) {
      long as, bs;
      const R* a = FN(pa)(blockA, depth, rows, i, &as);
      const R* b = FN(pb)(blockB, depth, cols, j, &bs);
      R acc = 0;
      for (long k = 0; k < depth; ++k) acc = FMA(a[k * as], b[k * bs], acc);
      { const volatile R t = alpha * acc; res[i + j * ldc] = res[i + j * ldc] + t; }
    }
}
#else
/* gebp_kernel, complex x complex (gebp_traits :590-744).  Two real accumulators per C element
 * ("DoublePacket", :566-571): first += a*Re(b), second += a*Im(b) on the (re,im) pair of a (:701-705), combined
 * with the conjugation pattern of traits.acc (:714-738), then r = tmp*alpha + r. */
static void FN(gebp)(R* res, long ldc, const R* blockA, const R* blockB, long rows, long depth, long cols,
                     const R* alpha, int conjl, int conjr) {
  for (long j = 0; j < cols; ++j)
    for (long i = 0; i < rows; ++i) {
      R f0 = 0, f1 = 0, s0 = 0, s1 = 0;
      long as, bs;
      const R* a0 = FN(pa)(blockA, depth, rows, i, &as);
      const R* b0 = FN(pb)(blockB, depth, cols, j, &bs);
      for (long k = 0; k < depth; ++k) {
        const R* a = a0 + k * as;
        const R* b = b0 + k * bs;
        f0 = FMA(a[0], b[0], f0); f1 = FMA(a[1], b[0], f1);
        s0 = FMA(a[0], b[1], s0); s1 = FMA(a[1], b[1], s1);
      }
      R t0, t1;
      if (!conjl && !conjr) { t0 = f0 - s1; t1 = f1 + s0; }
      else if (!conjl && conjr) { t0 = f0 + s1; t1 = f1 - s0; }
      else if (conjl && !conjr) { t0 = f0 + s1; t1 = -f1 + s0; }
      else { t0 = f0 - s1; t1 = -f1 - s0; }
      R* r = res + 2 * (i + j * ldc);
      /* pmadd(tmp, alpha, r) with complex pmul */
      /* Packet pmul for complex = mul, mul, addsub (arch/AVX/Complex.h); addsub is a builtin, so no FMA here */
      const volatile R e0 = t0 * alpha[0], e1 = t0 * alpha[1], o0 = t1 * alpha[1], o1 = t1 * alpha[0];
      const volatile R pr = e0 - o0, pi = e1 + o1;
      r[0] = pr + r[0]; r[1] = pi + r[1];
    }
}
#endif

/* general_matrix_matrix_product<...,ColMajor>::run, sequential branch (GeneralMatrixMatrix.h:155-198):
 * for i2 (mc) { for k2 (kc) { pack_lhs; for j2 (nc) { pack_rhs (once if pack_rhs_once); gebp } } }.
 * opa/opb as in blas/level3_impl.h:17-39: TR -> RowMajor source, ADJ -> RowMajor + Conj. */
static void FN(run)(long rows, long cols, long depth, const R* lhs, long lhsStride, int opa, const R* rhs,
                    long rhsStride, int opb, R* res, long resStride, const R* alpha, long kc, long mc, long nc) {
  if (mc > rows) mc = rows;
  if (nc > cols) nc = cols;
  if (kc > depth) kc = depth;
  R* blockA = (R*)malloc(sizeof(R) * NC * (size_t)(kc * mc));
  R* blockB = (R*)malloc(sizeof(R) * NC * (size_t)(kc * nc));
  const int ordA = opa == ORACLE_NOTR ? 0 : 1, ordB = opb == ORACLE_NOTR ? 0 : 1;
  const int pack_rhs_once = mc != rows && kc == depth && nc == cols;
  for (long i2 = 0; i2 < rows; i2 += mc) {
    const long amc = (i2 + mc < rows ? i2 + mc : rows) - i2;
    for (long k2 = 0; k2 < depth; k2 += kc) {
      const long akc = (k2 + kc < depth ? k2 + kc : depth) - k2;
      /* conjugation is NOT folded at pack time here: run() instantiates the packers with Conjugate=false and
       * hands ConjugateLhs/Rhs to gebp_kernel (GeneralMatrixMatrix.h:78-80) */
      FN(pack_lhs)(blockA, FN(at)(lhs, lhsStride, ordA, i2, k2), lhsStride, ordA, 0, akc, amc);
      for (long j2 = 0; j2 < cols; j2 += nc) {
        const long anc = (j2 + nc < cols ? j2 + nc : cols) - j2;
        if (!pack_rhs_once || i2 == 0)
          FN(pack_rhs)(blockB, FN(at)(rhs, rhsStride, ordB, k2, j2), rhsStride, ordB, 0, akc, anc);
#if NC == 1
        FN(gebp)(res + (i2 + j2 * resStride), resStride, blockA, blockB, amc, akc, anc, alpha[0]);
#else
        FN(gebp)(res + 2 * (i2 + j2 * resStride), resStride, blockA, blockB, amc, akc, anc, alpha,
                 opa == ORACLE_ADJ, opb == ORACLE_ADJ);
#endif
      }
    }
  }
  free(blockA);
  free(blockB);
}

/* beta pre-pass of blas/level3_impl.h:62-66: beta==0 -> setZero (C never read), beta!=1 -> C *= beta */
static void FN(scale_c)(R* c, long m, long n, long ldc, const R* beta) {
#if NC == 1
  if (beta[0] == (R)1) return;
  for (long j = 0; j < n; ++j)
    for (long i = 0; i < m; ++i) c[i + j * ldc] = beta[0] == 0 ? (R)0 : c[i + j * ldc] * beta[0];
#else
  if (beta[0] == (R)1 && beta[1] == 0) return;
  const int zero = beta[0] == 0 && beta[1] == 0;
  for (long j = 0; j < n; ++j)
    for (long i = 0; i < m; ++i) {
      R* z = c + 2 * (i + j * ldc);
      if (zero) { z[0] = 0; z[1] = 0; }
      else { const R re = z[0] * beta[0] - z[1] * beta[1], im = z[0] * beta[1] + z[1] * beta[0]; z[0] = re; z[1] = im; }
    }
#endif
}

/* EIGEN_BLAS_FUNC(gemm), blas/level3_impl.h:12-76: argument checks in positional order -> xerbla_, quick
 * returns, beta pre-pass, then the single-threaded blocked product (info = 0, :74). */
static int FN(blas_gemm)(const char* name, const char* opa, const char* opb, const int* m, const int* n, const int* k,
                         const R* alpha, const R* a, const int* lda, const R* b, const int* ldb, const R* beta, R* c,
                         const int* ldc, int threads) {
  const int oa = oracle_op(*opa), ob = oracle_op(*opb);
  int info = 0;
  if (oa == ORACLE_INVALID) info = 1;
  else if (ob == ORACLE_INVALID) info = 2;
  else if (*m < 0) info = 3;
  else if (*n < 0) info = 4;
  else if (*k < 0) info = 5;
  else if (*lda < imax(1, oa == ORACLE_NOTR ? *m : *k)) info = 8;
  else if (*ldb < imax(1, ob == ORACLE_NOTR ? *k : *n)) info = 10;
  else if (*ldc < imax(1, *m)) info = 13;
  if (info) return oracle_call_xerbla(name, &info);
  if (*m == 0 || *n == 0) return 0;
  FN(scale_c)(c, *m, *n, *ldc, beta);
  if (*k == 0) return 0;
  if (threads <= 1) {
    long kc = *k, mc = *m, nc = *n;
    oracle_blocking_sizes(TYPE_CODE, &kc, &mc, &nc, 1);
    FN(run)(*m, *n, *k, a, *lda, oa, b, *ldb, ob, c, *ldc, alpha, kc, mc, nc);
    return 0;
  }
  /* parallelize_gemm (Parallelizer.h:85-157) + the OpenMP branch of run (GeneralMatrixMatrix.h:83-152):
   * thread t owns the column slab [c0, c0+nc_t) of C; blocking is recomputed for T threads
   * (gemm_blocking_space::initParallel, GeneralMatrixMatrix.h:363-374); per kc panel every slab is a gebp over
   * each A' row slice.  Each C element still sees: per kc block one FMA chain, then one FMA with alpha. */
  long col0[256], ncols[256], row0[256], nrows[256];
  if (threads > 256) threads = 256;
  const int T = oracle_parallel_partition(TYPE_CODE, *m, *n, *k, threads, 0, col0, ncols, row0, nrows);
  if (T <= 1) {
    long kc = *k, mc = *m, nc = *n;
    oracle_blocking_sizes(TYPE_CODE, &kc, &mc, &nc, 1);
    FN(run)(*m, *n, *k, a, *lda, oa, b, *ldb, ob, c, *ldc, alpha, kc, mc, nc);
    return 0;
  }
  long kc = *k, mc = *m, nc = *n;
  oracle_blocking_sizes(TYPE_CODE, &kc, &mc, &nc, T);
  const int ordB = ob == ORACLE_NOTR ? 0 : 1;
#pragma omp parallel for num_threads(T) schedule(static, 1)
  for (int t = 0; t < T; ++t) {
    if (ncols[t] <= 0) continue;
    /* rows are not blocked in the OpenMP branch (mc = rows, :95); nc blocks the slab's own columns (:135-144) */
    FN(run)(*m, ncols[t], *k, a, *lda, oa, FN(at)(b, *ldb, ordB, 0, col0[t]), *ldb, ob,
            c + NC * ((long)col0[t] * *ldc), *ldc, alpha, kc, *m, nc);
  }
  return 0;
}

#undef PKT
#undef MR
#undef NR
#undef FN
#undef CAT
#undef CAT_
