// eigen_b200/csrc/gemm_simt.cu -- SIMT FFMA/DFMA GEMM for all four BLAS scalar types (sm_100a).
//
// Role (BASELINE.json north_star "SIMT-FFMA variant covers small or skinny shapes"):
// the variant for shapes where tensor tiles lose (tiny, skinny) and for operands the tensor variants cannot take
// (e.g. unaligned float panels); every op(A)/op(B) in {N,T,C}, any lda/ldb/ldc, any m,n,k >= 0.  It replaces, for
// those shapes, the reference's gemm_pack_lhs/rhs + gebp_kernel
// (Eigen/src/Core/products/GeneralBlockPanelKernel.h:858-2105): operands are "packed" into shared-memory panels
// As[k][m], Bs[k][n] (conjugation folded in at pack time like conj_if, BlasUtil.h:43-124) and multiplied by a
// register-tiled FMA micro-kernel; the epilogue fuses alpha and beta, which the reference applies as a separate
// pass over C (blas/level3_impl.h:62-66).
#include "common.cuh"
#include "scalar.cuh"

namespace b200 {
namespace {

// Thread block = 16 x 16 threads.  Each thread owns RM x RN chunks of VE x VE results, VE = 16 bytes / sizeof(T):
// chunk (cm, cn) covers rows cm*16*VE + tx*VE + [0,VE) and columns cn*16*VE + ty*VE + [0,VE).  With this layout
// every shared-memory read is one conflict-free 16-byte LDS per chunk and every C access is a 16-byte segment.
template <typename T, int RM, int RN, int BK>
struct SimtCfg {
  static constexpr int VE = 16 / (int)sizeof(T);
  static constexpr int TM = VE * RM, TN = VE * RN;
  static constexpr int BM = 16 * TM, BN = 16 * TN;
  static constexpr int LDSA = BM + VE, LDSB = BN + VE;  // padded row strides, multiples of 16 bytes
  static constexpr int EA = BM * BK / 256, EB = BN * BK / 256;  // elements fetched per thread and k-tile
  static_assert(BM * BK % 256 == 0 && BN * BK % 256 == 0, "tile must divide across 256 threads");
};

template <typename T, int RM, int RN, int BK>
__global__ void __launch_bounds__(256)
simt_gemm_kernel(int opa, int opb, int64_t m, int64_t n, int64_t k, T alpha, T beta, bool beta_zero,
                 const T* __restrict__ A, int64_t lda, const T* __restrict__ B, int64_t ldb, T* __restrict__ C,
                 int64_t ldc, int uplo, int herm) {
  using Cfg = SimtCfg<T, RM, RN, BK>;
  using S = Sc<T>;
  constexpr int VE = Cfg::VE, BM = Cfg::BM, BN = Cfg::BN, LDSA = Cfg::LDSA, LDSB = Cfg::LDSB;
  __shared__ __align__(16) T As[2][BK * LDSA];
  __shared__ __align__(16) T Bs[2][BK * LDSB];

  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.x * BM, n0 = (int64_t)blockIdx.y * BN;
  if (tile_outside(uplo, m0, m0 + BM, n0, n0 + BN)) return;   // rank-k update: tile entirely in the other triangle

  T acc[Cfg::TM][Cfg::TN];
#pragma unroll
  for (int i = 0; i < Cfg::TM; ++i)
#pragma unroll
    for (int j = 0; j < Cfg::TN; ++j) acc[i][j] = S::zero();

  T ra[Cfg::EA], rb[Cfg::EB];
  // "pack": element e of the A tile -> (i, kk); the thread->element map follows the contiguous direction of the
  // source (m for 'N', k for 'T'/'C') so that global loads coalesce.
  auto fetch = [&](int64_t k0) {
#pragma unroll
    for (int e = 0; e < Cfg::EA; ++e) {
      const int idx = tid + e * 256;
      int i, kk;
      if (opa == OP_N) { i = idx % BM; kk = idx / BM; } else { kk = idx % BK; i = idx / BK; }
      const int64_t gi = m0 + i, gk = k0 + kk;
      T v = S::zero();
      if (gi < m && gk < k) {
        v = (opa == OP_N) ? A[gi + gk * lda] : A[gk + gi * lda];
        if (opa == OP_C) v = S::conj(v);
      }
      ra[e] = v;
    }
#pragma unroll
    for (int e = 0; e < Cfg::EB; ++e) {
      const int idx = tid + e * 256;
      int j, kk;
      if (opb == OP_N) { kk = idx % BK; j = idx / BK; } else { j = idx % BN; kk = idx / BN; }
      const int64_t gj = n0 + j, gk = k0 + kk;
      T v = S::zero();
      if (gj < n && gk < k) {
        v = (opb == OP_N) ? B[gk + gj * ldb] : B[gj + gk * ldb];
        if (opb == OP_C) v = S::conj(v);
      }
      rb[e] = v;
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int e = 0; e < Cfg::EA; ++e) {
      const int idx = tid + e * 256;
      int i, kk;
      if (opa == OP_N) { i = idx % BM; kk = idx / BM; } else { kk = idx % BK; i = idx / BK; }
      As[buf][kk * LDSA + i] = ra[e];
    }
#pragma unroll
    for (int e = 0; e < Cfg::EB; ++e) {
      const int idx = tid + e * 256;
      int j, kk;
      if (opb == OP_N) { kk = idx % BK; j = idx / BK; } else { j = idx % BN; kk = idx / BN; }
      Bs[buf][kk * LDSB + j] = rb[e];
    }
  };

  const int64_t nkt = (k + BK - 1) / BK;
  if (nkt > 0) { fetch(0); stash(0); }
  __syncthreads();
  for (int64_t kt = 0; kt < nkt; ++kt) {
    const int buf = (int)(kt & 1);
    if (kt + 1 < nkt) fetch((kt + 1) * BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      T a[Cfg::TM], b[Cfg::TN];
#pragma unroll
      for (int c = 0; c < RM; ++c) {
        const int4 v = *reinterpret_cast<const int4*>(&As[buf][kk * LDSA + c * 16 * VE + tx * VE]);
        *reinterpret_cast<int4*>(&a[c * VE]) = v;
      }
#pragma unroll
      for (int c = 0; c < RN; ++c) {
        const int4 v = *reinterpret_cast<const int4*>(&Bs[buf][kk * LDSB + c * 16 * VE + ty * VE]);
        *reinterpret_cast<int4*>(&b[c * VE]) = v;
      }
#pragma unroll
      for (int i = 0; i < Cfg::TM; ++i)
#pragma unroll
        for (int j = 0; j < Cfg::TN; ++j) S::fma(acc[i][j], a[i], b[j]);
    }
    if (kt + 1 < nkt) stash(buf ^ 1);
    __syncthreads();
  }

  // fused epilogue: C = alpha*acc + beta*C; beta == 0 never reads C (blas/level3_impl.h:64).  Per column, all loads
  // are issued before the first store: a load that follows a store through the same pointer cannot be hoisted by the
  // compiler and every read-modify-write would pay a full memory round trip.
#pragma unroll
  for (int cn = 0; cn < RN; ++cn)
#pragma unroll
    for (int jn = 0; jn < VE; ++jn) {
      const int64_t gj = n0 + cn * 16 * VE + ty * VE + jn;
      if (gj >= n) continue;
      T old[Cfg::TM];
#pragma unroll
      for (int cm = 0; cm < RM; ++cm)
#pragma unroll
        for (int im = 0; im < VE; ++im) {
          const int64_t gi = m0 + cm * 16 * VE + tx * VE + im;
          old[cm * VE + im] = (!beta_zero && gi < m && in_triangle(uplo, gi, gj)) ? C[gi + gj * ldc] : S::zero();
        }
#pragma unroll
      for (int cm = 0; cm < RM; ++cm)
#pragma unroll
        for (int im = 0; im < VE; ++im) {
          const int64_t gi = m0 + cm * 16 * VE + tx * VE + im;
          if (gi >= m || !in_triangle(uplo, gi, gj)) continue;
          T r = S::mul(alpha, acc[cm * VE + im][cn * VE + jn]);
          if (!beta_zero) S::fma(r, beta, old[cm * VE + im]);
          if constexpr (sizeof(T) != sizeof(typename S::real)) { if (herm && gi == gj) r.y = 0; }
          C[gi + gj * ldc] = r;
        }
    }
}

template <typename T, int RM, int RN, int BK>
int launch_t(const GemmProblem& p, cudaStream_t s) {
  using Cfg = SimtCfg<T, RM, RN, BK>;
  using S = Sc<T>;
  if (p.m == 0 || p.n == 0) return 0;
  dim3 grid((unsigned)((p.m + Cfg::BM - 1) / Cfg::BM), (unsigned)((p.n + Cfg::BN - 1) / Cfg::BN));
  if (grid.y > 65535u) return (int)cudaErrorInvalidConfiguration;
  T alpha, beta;
  // k == 0: only the beta scaling happens (blas/level3_impl.h:68-69), whatever alpha is
  const bool no_product = (p.k == 0);
  if constexpr (sizeof(T) == sizeof(typename S::real)) {
    alpha = (T)(no_product ? 0.0 : p.alpha[0]);
    beta = (T)p.beta[0];
  } else {
    alpha.x = (typename S::real)(no_product ? 0.0 : p.alpha[0]);
    alpha.y = (typename S::real)(no_product ? 0.0 : p.alpha[1]);
    beta.x = (typename S::real)p.beta[0];
    beta.y = (typename S::real)p.beta[1];
  }
  const bool beta_zero = (p.beta[0] == 0.0 && p.beta[1] == 0.0);
  simt_gemm_kernel<T, RM, RN, BK><<<grid, 256, 0, s>>>(p.opa, p.opb, p.m, p.n, p.k, alpha, beta, beta_zero,
                                                       (const T*)p.A, p.lda, (const T*)p.B, p.ldb, (T*)p.C, p.ldc,
                                                       p.uplo, p.herm);
  count_launch();
  return (int)cudaGetLastError();
}

}  // namespace

int launch_simt(const GemmProblem& p, cudaStream_t s) {
  switch (p.type) {
    case TY_S: note_variant("simt_ffma_128x128x8"); return launch_t<float, 2, 2, 8>(p, s);
    case TY_D: note_variant("simt_dfma_64x64x8"); return launch_t<double, 2, 2, 8>(p, s);
    case TY_C: note_variant("simt_cffma_64x64x8"); return launch_t<float2, 2, 2, 8>(p, s);
    default: note_variant("simt_zdfma_32x32x8"); return launch_t<double2, 2, 2, 8>(p, s);
  }
}

}  // namespace b200
