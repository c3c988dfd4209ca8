// eigen_b200/csrc/gemm_dmma.cu -- FP64 tensor-core GEMM for double and complex<double> (sm_100a).
//
// Replaces the reference's gemm_pack_lhs/rhs + gebp_kernel pair for double / complex<double>
// (Eigen/src/Core/products/GeneralBlockPanelKernel.h:858-2105) and the blocked driver
// general_matrix_matrix_product::run (GeneralMatrixMatrix.h:59-199):
//   * "packing" = a 4-stage cp.async pipeline that lays the A and B panels of one tile of C into shared memory as
//     k-major panels As[k][m], Bs[k][n] (row stride = 4 mod 16 doubles => every fragment LDS.64 is bank-conflict
//     free); 8-byte cp.async handles any lda/ldb and any 8-byte aligned pointer, 16-byte cp.async is used when the
//     panel is contiguous along the tile dimension and 16-byte aligned.  Per-thread source pointers are computed
//     once and only advanced in the k loop, and the copies of tile kt+3 are interleaved with the MMAs of tile kt;
//   * "gebp" = warp-level mma.sync.aligned.m8n8k4.f64 (SASS DMMA.8x8x4, the native FP64 tensor op on sm_100a;
//     tcgen05 has no f64 kind), FP64 accumulators in registers;
//   * complex<double> reuses the same main loop on the interleaved real view (the reference's DoublePacket trick,
//     GeneralBlockPanelKernel.h:566-571,701-740): Ahat (2m x k) holds (re,im) rows, Bhat (k x 2n) holds (re,im)
//     columns, P=Ar.Br, Q=Ar.Bi, R=Ai.Br, S=Ai.Bi; the epilogue combines re = P -/+ S, im = +/-Q +/- R with the four
//     conjugation sign patterns of gebp_traits::acc (:714-738);
//   * the epilogue fuses alpha and beta (the reference scales C by beta in a separate pass, blas/level3_impl.h:62-66).
// Tile configurations (template Cfg): 128x128 tile / 8 warps of 64x32 / 1 CTA per SM, 128x64 tile / 8 warps of
// 32x32 / 2 CTAs per SM, 128x128 tile / 16 warps of 32x32; the default is chosen from measurements
// (profiles/dmma_cfg_sweep_r01.md), B200BLAS_DMMA_CFG overrides it.
#include <cstdlib>

#include "common.cuh"

namespace b200 {
namespace {

template <int BM_, int BN_, int WM_, int WN_, int MINB_, int BK_ = 16, int STAGES_ = 4>
struct Cfg {
  static constexpr int BM = BM_, BN = BN_, WM = WM_, WN = WN_, MINB = MINB_, BK = BK_, STAGES = STAGES_;
  static constexpr int WARPS_M = BM / WM, WARPS_N = BN / WN;
  static constexpr int THREADS = WARPS_M * WARPS_N * 32;
  static constexpr int MI = WM / 8, NJ = WN / 8;
  static constexpr int LDA_S = BM + 4, LDB_S = BN + 4;  // doubles; both = 4 mod 16
  static constexpr int PANEL_A = BK * LDA_S, PANEL_B = BK * LDB_S;
  static constexpr int SMEM_BYTES = STAGES * (PANEL_A + PANEL_B) * (int)sizeof(double);
  static_assert(LDA_S % 16 == 4 && LDB_S % 16 == 4, "conflict-free fragment reads need stride = 4 mod 16");
};

struct PanelSrc {
  const double* base;  // real view of the operand
  int64_t ld;          // leading dimension in scalars (complex counts as one)
  int64_t dim;         // extent along the tile dimension in scalars (m for A, n for B)
  int dim_contig;      // 1: memory is contiguous along the tile dimension (A:'N', B:'T'/'C'), 0: along k
  int vec16;           // 1: 16-byte cp.async (alignment checked on the host); always 1 for complex
};

// Per-thread loader state for one operand panel of PR real rows: copy e (0 <= e < E) moves `bytes` bytes from
// p + e*estep to smem offset soff + e*sstep and belongs to k index kk0 + e*kkstep of the tile.
template <int PR, int THREADS, bool CPLX, int BK>
struct Loader {
  const double* p;
  uint32_t soff;
  int kk0;
  uint32_t vmask;   // bit e: the row/column of copy e is inside the matrix
  int vbytes;       // bytes actually read when valid (8 for the odd last row of a 16-byte real granule)
  // uniform (same for all threads)
  int64_t estep, kstep;
  int sstep, kkstep, bytes, E;

  __device__ __forceinline__ void init(const PanelSrc& s, int64_t dim0, int tid, int lds_row) {
    const int gran = (CPLX || s.vec16) ? 2 : 1;        // doubles per copy
    const int rs = CPLX ? 2 : 1;                        // doubles per scalar
    const int GD = PR / gran;                           // granules along the tile dimension
    bytes = 8 * gran;
    E = GD * BK / THREADS;
    vmask = 0;
    vbytes = bytes;
    if (s.dim_contig) {
      const int rg = tid % GD;
      kk0 = tid / GD;
      kkstep = THREADS / GD;
      const int64_t gd = dim0 + (int64_t)rg * gran / rs;  // scalar index along the tile dimension
      p = s.base + rs * gd + (int64_t)kk0 * s.ld * rs;
      estep = (int64_t)kkstep * s.ld * rs;
      kstep = (int64_t)BK * s.ld * rs;
      soff = (uint32_t)(kk0 * lds_row + rg * gran);
      sstep = kkstep * lds_row;
      if (gd < s.dim) {  // gd = first scalar of the granule
        vmask = 0xffffffffu;
        if (!CPLX && gran == 2 && gd + 1 >= s.dim) vbytes = 8;
      }
    } else {
      kk0 = tid % BK;
      kkstep = 0;
      const int r0 = tid / BK;                          // scalar index inside the tile for e = 0
      const int rstep = THREADS / BK;
      p = s.base + rs * ((int64_t)kk0 + (dim0 + r0) * s.ld);
      estep = (int64_t)rstep * s.ld * rs;
      kstep = (int64_t)BK * rs;
      soff = (uint32_t)(kk0 * lds_row + r0 * rs);
      sstep = rstep * rs;
#pragma unroll
      for (int e = 0; e < 16; ++e)
        if (e < E && dim0 + r0 + (int64_t)e * rstep < s.dim) vmask |= 1u << e;
    }
  }
  // issue copy e of one k-tile: ptile = p advanced to the tile (computed once per tile), krem = k - k0 clamped to int
  __device__ __forceinline__ void copy(double* S, const double* fallback, int e, const double* ptile, int krem) const {
    const bool ok = ((vmask >> e) & 1u) && (kk0 + e * kkstep < krem);
    const double* src = ok ? ptile + (int64_t)e * estep : fallback;
    double* dst = S + soff + e * sstep;
    if (bytes == 16) cp_async_zfill<16>(dst, src, ok ? vbytes : 0);
    else cp_async_zfill<8>(dst, src, ok ? 8 : 0);
  }
  __device__ __forceinline__ const double* tile_ptr(int64_t kt) const { return p + kt * kstep; }
};

__device__ __forceinline__ void dmma_8x8x4(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

struct EpiParams {
  double alpha[2], beta[2];
  int beta_zero;
  int conja, conjb;
};

// grouped tile order: consecutive CTAs walk GROUP tile-rows first so that one wave of CTAs shares a compact
// block of A and B panels in L2.
constexpr int GROUP = 16;
__device__ __forceinline__ void tile_of(int64_t pid, int64_t tiles_m, int64_t tiles_n, int64_t& tm, int64_t& tn) {
  const int64_t per_group = GROUP * tiles_n;
  const int64_t g = pid / per_group;
  const int64_t first = g * GROUP;
  const int64_t gsz = (tiles_m - first < GROUP) ? (tiles_m - first) : GROUP;
  const int64_t in = pid - g * per_group;
  tm = first + in % gsz;
  tn = in / gsz;
}

template <typename C, bool CPLX>
__global__ void __launch_bounds__(C::THREADS, C::MINB)
dmma_gemm_kernel(PanelSrc a, PanelSrc b, int64_t m, int64_t n, int64_t k, double* __restrict__ Cmat, int64_t ldc,
                 EpiParams ep, int64_t tiles_m, int64_t tiles_n) {
  extern __shared__ __align__(16) double smem[];
  double* As = smem;
  double* Bs = smem + C::STAGES * C::PANEL_A;
  constexpr int MI = C::MI, NJ = C::NJ;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp % C::WARPS_M, wn = warp / C::WARPS_M;
  int64_t tm, tn;
  tile_of(blockIdx.x, tiles_m, tiles_n, tm, tn);
  constexpr int SC = CPLX ? 2 : 1;  // real rows/cols per scalar
  const int64_t m0 = tm * (C::BM / SC), n0 = tn * (C::BN / SC);  // tile origin in scalars

  constexpr int BK = C::BK, STAGES = C::STAGES;
  Loader<C::BM, C::THREADS, CPLX, BK> la;
  Loader<C::BN, C::THREADS, CPLX, BK> lb;
  la.init(a, m0, tid, C::LDA_S);
  lb.init(b, n0, tid, C::LDB_S);

  double acc[MI][NJ][2];
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NJ; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

  const int64_t nkt = (k + BK - 1) / BK;
  auto load_all = [&](int64_t kt) {
    double* sa = As + (kt % STAGES) * C::PANEL_A;
    double* sb = Bs + (kt % STAGES) * C::PANEL_B;
#pragma unroll
    const int64_t rem = k - kt * BK;
    const int krem = rem > (1 << 20) ? (1 << 20) : (int)rem;
    const double* pa = la.tile_ptr(kt);
    const double* pb = lb.tile_ptr(kt);
#pragma unroll
    for (int e = 0; e < 16; ++e) if (e < la.E) la.copy(sa, a.base, e, pa, krem);
#pragma unroll
    for (int e = 0; e < 16; ++e) if (e < lb.E) lb.copy(sb, b.base, e, pb, krem);
  };
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nkt) load_all(s);
    cp_async_commit();
  }

  const int fr = lane >> 2, fk = lane & 3;  // fragment row (0..7) and k (0..3) of this lane
  for (int64_t kt = 0; kt < nkt; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    const int64_t nxt = kt + STAGES - 1;
    const bool do_load = nxt < nkt;
    double* sa_n = As + (nxt % STAGES) * C::PANEL_A;
    double* sb_n = Bs + (nxt % STAGES) * C::PANEL_B;
    const int64_t rem_n = k - nxt * BK;
    const int krem_n = rem_n > (1 << 20) ? (1 << 20) : (int)rem_n;
    const double* pa_n = la.tile_ptr(nxt);
    const double* pb_n = lb.tile_ptr(nxt);
    const double* As_ = As + (kt % STAGES) * C::PANEL_A + wm * C::WM + fr;
    const double* Bs_ = Bs + (kt % STAGES) * C::PANEL_B + wn * C::WN + fr;
#pragma unroll
    for (int k4 = 0; k4 < BK / 4; ++k4) {
      double af[MI], bf[NJ];
#pragma unroll
      for (int i = 0; i < MI; ++i) af[i] = As_[(k4 * 4 + fk) * C::LDA_S + i * 8];
#pragma unroll
      for (int j = 0; j < NJ; ++j) bf[j] = Bs_[(k4 * 4 + fk) * C::LDB_S + j * 8];
      // a quarter of the next tile's copies rides along with each k4 step
      if (do_load) {
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          if ((e % (BK / 4)) == k4) {
            if (e < la.E) la.copy(sa_n, a.base, e, pa_n, krem_n);
            if (e < lb.E) lb.copy(sb_n, b.base, e, pb_n, krem_n);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) dmma_8x8x4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
    cp_async_commit();
  }
  cp_async_wait<0>();

  // ---- epilogue: registers -> global, alpha/beta fused ------------------------------------------------------
  // accumulator (i,j): real row wm*WM + i*8 + lane/4, real cols wn*WN + j*8 + 2*(lane%4) + {0,1}
  if constexpr (!CPLX) {
    const double alpha = ep.alpha[0], beta = ep.beta[0];
#pragma unroll
    for (int j = 0; j < NJ; ++j)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int64_t gj = n0 + wn * C::WN + j * 8 + 2 * fk + c;
        if (gj >= n) continue;
#pragma unroll
        for (int i = 0; i < MI; ++i) {
          const int64_t gi = m0 + wm * C::WM + i * 8 + fr;
          if (gi >= m) continue;
          double* pc = Cmat + gi + gj * ldc;
          double r = alpha * acc[i][j][c];
          if (!ep.beta_zero) r = fma(beta, *pc, r);
          *pc = r;
        }
      }
  } else {
    // lanes (l, l^4) hold the Ar-row and the Ai-row of one complex row; c0 = x.Br, c1 = x.Bi
    const bool odd = fr & 1;  // this lane holds the Ai row
    // re = P + sS*S, im = sQ*Q + sR*R  (gebp_traits::acc sign patterns, GeneralBlockPanelKernel.h:714-738)
    const double sS = (ep.conja != ep.conjb) ? 1.0 : -1.0;
    const double sQ = ep.conjb ? -1.0 : 1.0;
    const double sR = ep.conja ? -1.0 : 1.0;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int64_t gj = n0 + wn * (C::WN / 2) + j * 4 + fk;  // complex column
#pragma unroll
      for (int i = 0; i < MI; ++i) {
        const int64_t gi = m0 + wm * (C::WM / 2) + i * 4 + (fr >> 1);  // complex row
        const double c0 = acc[i][j][0], c1 = acc[i][j][1];
        const double o1 = __shfl_xor_sync(0xffffffffu, c1, 4);
        // even lane: re = P + sS*S (P = own c0, S = partner c1); odd lane: im = sQ*Q + sR*R (Q = partner c1, R = own c0)
        const double mine = odd ? fma(sQ, o1, sR * c0) : fma(sS, o1, c0);
        const double other = __shfl_xor_sync(0xffffffffu, mine, 4);
        const double re = odd ? other : mine, im = odd ? mine : other;
        if (gi >= m || gj >= n) continue;
        double* pc = Cmat + 2 * (gi + gj * ldc);
        // out = alpha*(re,im) + beta*Cold ; even lane stores the real part, odd lane the imaginary part
        double out = odd ? fma(ep.alpha[0], im, ep.alpha[1] * re) : fma(ep.alpha[0], re, -ep.alpha[1] * im);
        if (!ep.beta_zero) {
          const double cr = pc[0], ci = pc[1];
          out += odd ? fma(ep.beta[0], ci, ep.beta[1] * cr) : fma(ep.beta[0], cr, -ep.beta[1] * ci);
        }
        pc[odd ? 1 : 0] = out;
      }
    }
  }
}

bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

using CfgA = Cfg<128, 128, 64, 32, 1>;  // 8 warps, 64x32 warp tiles, 1 CTA / SM
using CfgB = Cfg<128, 64, 32, 32, 2>;   // 8 warps, 32x32 warp tiles, 2 CTAs / SM
using CfgC = Cfg<128, 128, 32, 32, 1>;  // 16 warps, 32x32 warp tiles, 1 CTA / SM
using CfgD = Cfg<128, 64, 32, 32, 2, 32, 2>;  // as B with BK = 32, double buffered: half as many CTA barriers

template <typename C, bool CPLX>
int launch_cfg(const GemmProblem& p, cudaStream_t s, const PanelSrc& a, const PanelSrc& b, const EpiParams& ep) {
  const int sc = CPLX ? 2 : 1;
  const int64_t tiles_m = (p.m * sc + C::BM - 1) / C::BM, tiles_n = (p.n * sc + C::BN - 1) / C::BN;
  const int64_t tiles = tiles_m * tiles_n;
  if (tiles > 0x7fffffffLL) return (int)cudaErrorInvalidConfiguration;
  static bool attr_done = false;
  if (!attr_done) {
    B200_CUDA_TRY(cudaFuncSetAttribute(dmma_gemm_kernel<C, CPLX>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_done = true;
  }
  dmma_gemm_kernel<C, CPLX><<<(unsigned)tiles, C::THREADS, C::SMEM_BYTES, s>>>(a, b, p.m, p.n, p.k, (double*)p.C, p.ldc, ep,
                                                                             tiles_m, tiles_n);
  count_launch();
  return (int)cudaGetLastError();
}

int dmma_cfg() {
  static int cfg = [] {
    const char* e = getenv("B200BLAS_DMMA_CFG");
    if (!e) return -1;
    return (e[0] == 'A' || e[0] == 'a') ? 0 : (e[0] == 'B' || e[0] == 'b') ? 1 : (e[0] == 'C' || e[0] == 'c') ? 2
           : (e[0] == 'D' || e[0] == 'd') ? 3 : -1;
  }();
  return cfg;
}

}  // namespace

bool dmma_supported(const GemmProblem& p) {
  if (p.type != TY_D && p.type != TY_Z) return false;
  if (((uintptr_t)p.A & 7) || ((uintptr_t)p.B & 7) || ((uintptr_t)p.C & 7)) return false;
  if (p.type == TY_Z && (!aligned16(p.A) || !aligned16(p.B))) return false;  // 16-byte cp.async per complex scalar
  return p.m > 0 && p.n > 0 && p.k > 0;
}

int launch_dmma(const GemmProblem& p, cudaStream_t s) {
  const bool cplx = p.type == TY_Z;
  PanelSrc a, b;
  a.base = (const double*)p.A; a.ld = p.lda; a.dim = p.m; a.dim_contig = (p.opa == OP_N);
  b.base = (const double*)p.B; b.ld = p.ldb; b.dim = p.n; b.dim_contig = (p.opb != OP_N);
  // 16-byte cp.async: real panels need a contiguous tile dimension, an even ld and a 16-byte aligned base;
  // complex scalars are 16 bytes themselves (alignment checked in dmma_supported).
  a.vec16 = cplx ? 1 : (a.dim_contig && aligned16(p.A) && (p.lda % 2 == 0));
  b.vec16 = cplx ? 1 : (b.dim_contig && aligned16(p.B) && (p.ldb % 2 == 0));
  EpiParams ep;
  ep.alpha[0] = p.alpha[0]; ep.alpha[1] = p.alpha[1];
  ep.beta[0] = p.beta[0]; ep.beta[1] = p.beta[1];
  ep.beta_zero = (p.beta[0] == 0.0 && p.beta[1] == 0.0);
  ep.conja = (p.opa == OP_C); ep.conjb = (p.opb == OP_C);
  int cfg = dmma_cfg();
  if (cfg < 0) cfg = 1;  // default: 128x64 tiles, 2 CTAs / SM (profiles/dmma_cfg_sweep_r01.md)
  if (cplx) {
    switch (cfg) {
      case 0: note_variant("dmma_z_A_64x64x16_w32x16"); return launch_cfg<CfgA, true>(p, s, a, b, ep);
      case 2: note_variant("dmma_z_C_64x64x16_w16x16"); return launch_cfg<CfgC, true>(p, s, a, b, ep);
      case 3: note_variant("dmma_z_D_64x32x32_w16x16_2cta"); return launch_cfg<CfgD, true>(p, s, a, b, ep);
      default: note_variant("dmma_z_B_64x32x16_w16x16_2cta"); return launch_cfg<CfgB, true>(p, s, a, b, ep);
    }
  }
  switch (cfg) {
    case 0: note_variant("dmma_d_A_128x128x16_w64x32"); return launch_cfg<CfgA, false>(p, s, a, b, ep);
    case 2: note_variant("dmma_d_C_128x128x16_w32x32"); return launch_cfg<CfgC, false>(p, s, a, b, ep);
    case 3: note_variant("dmma_d_D_128x64x32_w32x32_2cta"); return launch_cfg<CfgD, false>(p, s, a, b, ep);
    default: note_variant("dmma_d_B_128x64x16_w32x32_2cta"); return launch_cfg<CfgB, false>(p, s, a, b, ep);
  }
}

}  // namespace b200
