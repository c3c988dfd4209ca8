// oracle/ref_eigen_shim.cpp -- TEST INFRASTRUCTURE, not product code.
//
// A C-ABI window onto the UNMODIFIED reference (Eigen headers under $(REF), compiled where they lie by
// oracle/Makefile into oracle/_ref/libeigen_gebp_omp.so).  It lets the Python tests and bench.py's
// cpu_baseline / --impl reference legs drive the reference's own hot path:
//
//   * ref_eigen_gemm_{s,d,c,z}: C = beta*C ; C.noalias() += alpha*op(A)*op(B) through Eigen's public
//     expression API, i.e. generic_product_impl<GemmProduct>::scaleAndAddTo -> parallelize_gemm (OpenMP)
//     -> general_matrix_matrix_product::run -> gemm_pack_lhs/rhs + gebp_kernel
//     (Eigen/src/Core/products/GeneralMatrixMatrix.h:405-492, Parallelizer.h:85-157).
//     This is exactly what bench/bench_gemm.cpp:130-133 times.
//   * ref_blocking_sizes_*: computeProductBlockingSizes (GeneralBlockPanelKernel.h:296-308).
//   * ref_pack_lhs_* / ref_pack_rhs_*: gemm_pack_lhs / gemm_pack_rhs (GeneralBlockPanelKernel.h:1688-2105),
//     used to pin the oracle's packed-panel layouts byte for byte.
//
// Nothing here re-implements reference logic; it only instantiates and calls it.
#include <Eigen/Core>
#include <complex>

using namespace Eigen;

namespace {

template <typename S>
void gemm_expr(char ta, char tb, int m, int n, int k, S alpha, const S* a, int lda, const S* b, int ldb, S beta,
               S* c, int ldc, int nthreads) {
  typedef Matrix<S, Dynamic, Dynamic, ColMajor> Mat;
  typedef Map<const Mat, 0, OuterStride<> > CMap;
  typedef Map<Mat, 0, OuterStride<> > MMap;
  if (nthreads > 0) Eigen::setNbThreads(nthreads);
  const bool na = (ta == 'N' || ta == 'n'), nb = (tb == 'N' || tb == 'n');
  const bool ca = (ta == 'C' || ta == 'c'), cb = (tb == 'C' || tb == 'c');
  CMap A(a, na ? m : k, na ? k : m, OuterStride<>(lda));
  CMap B(b, nb ? k : n, nb ? n : k, OuterStride<>(ldb));
  MMap C(c, m, n, OuterStride<>(ldc));
  if (m == 0 || n == 0) return;
  if (beta != S(1)) {
    if (beta == S(0)) C.setZero(); else C *= beta;
  }
  if (k == 0) return;
  const int code = (na ? 0 : ca ? 2 : 1) + 3 * (nb ? 0 : cb ? 2 : 1);
  switch (code) {
    case 0: C.noalias() += alpha * (A * B); break;
    case 1: C.noalias() += alpha * (A.transpose() * B); break;
    case 2: C.noalias() += alpha * (A.adjoint() * B); break;
    case 3: C.noalias() += alpha * (A * B.transpose()); break;
    case 4: C.noalias() += alpha * (A.transpose() * B.transpose()); break;
    case 5: C.noalias() += alpha * (A.adjoint() * B.transpose()); break;
    case 6: C.noalias() += alpha * (A * B.adjoint()); break;
    case 7: C.noalias() += alpha * (A.transpose() * B.adjoint()); break;
    case 8: C.noalias() += alpha * (A.adjoint() * B.adjoint()); break;
  }
}

template <typename S>
void blocking(long* k, long* m, long* n, int threads) {
  Index kk = *k, mm = *m, nn = *n;
  internal::computeProductBlockingSizes<S, S>(kk, mm, nn, Index(threads));
  *k = kk; *m = mm; *n = nn;
}

template <typename S, int Order, bool Conj>
void pack_lhs(S* blockA, const S* lhs, long stride, long depth, long rows) {
  typedef internal::gebp_traits<S, S> Traits;
  typedef internal::const_blas_data_mapper<S, Index, Order> Mapper;
  internal::gemm_pack_lhs<S, Index, Mapper, Traits::mr, Traits::LhsProgress, Order, Conj> pack;
  pack(blockA, Mapper(lhs, stride), depth, rows);
}
template <typename S, int Order, bool Conj>
void pack_rhs(S* blockB, const S* rhs, long stride, long depth, long cols) {
  typedef internal::gebp_traits<S, S> Traits;
  typedef internal::const_blas_data_mapper<S, Index, Order> Mapper;
  internal::gemm_pack_rhs<S, Index, Mapper, Traits::nr, Order, Conj> pack;
  pack(blockB, Mapper(rhs, stride), depth, cols);
}

}  // namespace

extern "C" {

void ref_eigen_gemm_s(char ta, char tb, int m, int n, int k, const float* alpha, const float* a, int lda,
                      const float* b, int ldb, const float* beta, float* c, int ldc, int nthreads) {
  gemm_expr<float>(ta, tb, m, n, k, *alpha, a, lda, b, ldb, *beta, c, ldc, nthreads);
}
void ref_eigen_gemm_d(char ta, char tb, int m, int n, int k, const double* alpha, const double* a, int lda,
                      const double* b, int ldb, const double* beta, double* c, int ldc, int nthreads) {
  gemm_expr<double>(ta, tb, m, n, k, *alpha, a, lda, b, ldb, *beta, c, ldc, nthreads);
}
void ref_eigen_gemm_c(char ta, char tb, int m, int n, int k, const float* alpha, const float* a, int lda,
                      const float* b, int ldb, const float* beta, float* c, int ldc, int nthreads) {
  typedef std::complex<float> S;
  gemm_expr<S>(ta, tb, m, n, k, *(const S*)alpha, (const S*)a, lda, (const S*)b, ldb, *(const S*)beta, (S*)c, ldc,
               nthreads);
}
void ref_eigen_gemm_z(char ta, char tb, int m, int n, int k, const double* alpha, const double* a, int lda,
                      const double* b, int ldb, const double* beta, double* c, int ldc, int nthreads) {
  typedef std::complex<double> S;
  gemm_expr<S>(ta, tb, m, n, k, *(const S*)alpha, (const S*)a, lda, (const S*)b, ldb, *(const S*)beta, (S*)c, ldc,
               nthreads);
}

int ref_nb_threads(void) { return Eigen::nbThreads(); }
void ref_cache_sizes(long* l1, long* l2, long* l3) {
  *l1 = Eigen::l1CacheSize(); *l2 = Eigen::l2CacheSize(); *l3 = Eigen::l3CacheSize();
}
// register-block geometry gebp_traits<S,S>::{mr,nr,LhsProgress}; type: 0=s 1=d 2=c 3=z
void ref_gebp_traits(int type, int* mr, int* nr, int* lhs_progress) {
  switch (type) {
#define T_(S) { typedef internal::gebp_traits<S, S> T; *mr = T::mr; *nr = T::nr; *lhs_progress = T::LhsProgress; } break
    case 0: T_(float);
    case 1: T_(double);
    case 2: T_(std::complex<float>);
    default: T_(std::complex<double>);
#undef T_
  }
}
void ref_blocking_sizes(int type, long* k, long* m, long* n, int threads) {
  switch (type) {
    case 0: blocking<float>(k, m, n, threads); break;
    case 1: blocking<double>(k, m, n, threads); break;
    case 2: blocking<std::complex<float> >(k, m, n, threads); break;
    default: blocking<std::complex<double> >(k, m, n, threads); break;
  }
}

// order: 0 = ColMajor source, 1 = RowMajor source.  Output buffers must hold rows*depth (cols*depth) scalars.
void ref_pack_lhs_d(double* out, const double* lhs, long stride, long depth, long rows, int order) {
  if (order == 0) pack_lhs<double, ColMajor, false>(out, lhs, stride, depth, rows);
  else pack_lhs<double, RowMajor, false>(out, lhs, stride, depth, rows);
}
void ref_pack_rhs_d(double* out, const double* rhs, long stride, long depth, long cols, int order) {
  if (order == 0) pack_rhs<double, ColMajor, false>(out, rhs, stride, depth, cols);
  else pack_rhs<double, RowMajor, false>(out, rhs, stride, depth, cols);
}
void ref_pack_lhs_s(float* out, const float* lhs, long stride, long depth, long rows, int order) {
  if (order == 0) pack_lhs<float, ColMajor, false>(out, lhs, stride, depth, rows);
  else pack_lhs<float, RowMajor, false>(out, lhs, stride, depth, rows);
}
void ref_pack_rhs_s(float* out, const float* rhs, long stride, long depth, long cols, int order) {
  if (order == 0) pack_rhs<float, ColMajor, false>(out, rhs, stride, depth, cols);
  else pack_rhs<float, RowMajor, false>(out, rhs, stride, depth, cols);
}
void ref_pack_lhs_z(double* out, const double* lhs, long stride, long depth, long rows, int order, int conj) {
  typedef std::complex<double> S;
  if (order == 0) { if (conj) pack_lhs<S, ColMajor, true>((S*)out, (const S*)lhs, stride, depth, rows); else pack_lhs<S, ColMajor, false>((S*)out, (const S*)lhs, stride, depth, rows); }
  else { if (conj) pack_lhs<S, RowMajor, true>((S*)out, (const S*)lhs, stride, depth, rows); else pack_lhs<S, RowMajor, false>((S*)out, (const S*)lhs, stride, depth, rows); }
}
void ref_pack_rhs_z(double* out, const double* rhs, long stride, long depth, long cols, int order, int conj) {
  typedef std::complex<double> S;
  if (order == 0) { if (conj) pack_rhs<S, ColMajor, true>((S*)out, (const S*)rhs, stride, depth, cols); else pack_rhs<S, ColMajor, false>((S*)out, (const S*)rhs, stride, depth, cols); }
  else { if (conj) pack_rhs<S, RowMajor, true>((S*)out, (const S*)rhs, stride, depth, cols); else pack_rhs<S, RowMajor, false>((S*)out, (const S*)rhs, stride, depth, cols); }
}

}  // extern "C"
