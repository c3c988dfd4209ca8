// eigen_b200/csrc/gemm_dmma.cu -- FP64 tensor-core GEMM for double and complex<double> (sm_100a).
//
// Replaces the reference's gemm_pack_lhs/rhs + gebp_kernel pair for double / complex<double>
// (Eigen/src/Core/products/GeneralBlockPanelKernel.h:858-2105) and the blocked driver
// general_matrix_matrix_product::run (GeneralMatrixMatrix.h:59-199):
//   * "packing" = a 4-stage cp.async pipeline that lays the A and B panels of a 128x128 tile of C into shared
//     memory as k-major panels As[k][m], Bs[k][n] (row stride 132 doubles => every fragment LDS.64 is
//     bank-conflict free); 8-byte cp.async handles any lda/ldb and any 8-byte aligned pointer, 16-byte cp.async is
//     used when the panel is contiguous along the tile dimension and 16-byte aligned;
//   * "gebp" = warp-level mma.sync.aligned.m8n8k4.f64 (SASS DMMA.8x8x4, the native FP64 tensor op on sm_100a;
//     tcgen05 has no f64 kind).  8 warps x (64x32) warp tiles, 64 FP64 accumulators per thread;
//   * complex<double> reuses the same main loop on the interleaved real view (the reference's DoublePacket trick,
//     GeneralBlockPanelKernel.h:566-571,701-740): Ahat (2m x k) holds (re,im) rows, Bhat (k x 2n) holds (re,im)
//     columns, P=Ar.Br, Q=Ar.Bi, R=Ai.Br, S=Ai.Bi; the epilogue combines re = P -/+ S, im = +/-Q +/- R with the four
//     conjugation sign patterns of gebp_traits::acc (:714-738);
//   * the epilogue fuses alpha and beta (the reference scales C by beta in a separate pass, blas/level3_impl.h:62-66).
#include "common.cuh"

namespace b200 {
namespace {

constexpr int BM = 128, BN = 128, BK = 16, STAGES = 4, THREADS = 256;
constexpr int LDS_ROW = 132;                    // doubles per k-row of a panel; 132 % 16 == 4 => conflict-free fragments
constexpr int PANEL = BK * LDS_ROW;             // doubles per panel per stage
constexpr int SMEM_BYTES = STAGES * 2 * PANEL * (int)sizeof(double);  // 135168

struct PanelSrc {
  const double* base;  // real view of the operand
  int64_t ld;          // leading dimension in scalars (complex counts as one)
  int64_t dim;         // extent along the tile dimension in scalars (m for A, n for B)
  int dim_contig;      // 1: memory is contiguous along the tile dimension (A:'N', B:'T'/'C'), 0: along k
  int vec16;           // 1: 16-byte cp.async allowed (alignment checked on the host)
};

// Fill one k-major panel S[kk][r] (r = real row/column inside the tile, 0..127) for k in [k0, k0+16).
template <bool CPLX>
__device__ __forceinline__ void load_panel(double* S, const PanelSrc& src, int64_t dim0 /*scalars*/, int64_t k0,
                                           int64_t k, int tid) {
  if constexpr (!CPLX) {
    if (src.dim_contig && src.vec16) {
      // 16-byte granules: rows (r, r+1) of one k
#pragma unroll
      for (int e = 0; e < (BM * BK / 2) / THREADS; ++e) {
        const int g = tid + e * THREADS;
        const int r = (g % (BM / 2)) * 2, kk = g / (BM / 2);
        const int64_t gd = dim0 + r, gk = k0 + kk;
        int valid = 0;
        if (gk < k && gd < src.dim) valid = (src.dim - gd >= 2) ? 16 : 8;
        const double* p = valid ? src.base + gd + gk * src.ld : src.base;
        cp_async_zfill<16>(S + kk * LDS_ROW + r, p, valid);
      }
    } else {
#pragma unroll
      for (int e = 0; e < (BM * BK) / THREADS; ++e) {
        const int g = tid + e * THREADS;
        int r, kk;
        if (src.dim_contig) { r = g % BM; kk = g / BM; } else { kk = g % BK; r = g / BK; }
        const int64_t gd = dim0 + r, gk = k0 + kk;
        const bool ok = gk < k && gd < src.dim;
        const double* p = ok ? (src.dim_contig ? src.base + gd + gk * src.ld : src.base + gk + gd * src.ld) : src.base;
        cp_async_zfill<8>(S + kk * LDS_ROW + r, p, ok ? 8 : 0);
      }
    }
  } else {
    // one granule = one complex scalar -> real rows (2c, 2c+1) of the panel
#pragma unroll
    for (int e = 0; e < (BM / 2 * BK) / THREADS; ++e) {
      const int g = tid + e * THREADS;
      int c, kk;
      if (src.dim_contig) { c = g % (BM / 2); kk = g / (BM / 2); } else { kk = g % BK; c = g / BK; }
      const int64_t gd = dim0 + c, gk = k0 + kk;
      const bool ok = gk < k && gd < src.dim;
      const double* p = ok ? src.base + 2 * (src.dim_contig ? gd + gk * src.ld : gk + gd * src.ld) : src.base;
      double* d = S + kk * LDS_ROW + 2 * c;
      if (src.vec16) cp_async_zfill<16>(d, p, ok ? 16 : 0);
      else { cp_async_zfill<8>(d, p, ok ? 8 : 0); cp_async_zfill<8>(d + 1, ok ? p + 1 : p, ok ? 8 : 0); }
    }
  }
}

__device__ __forceinline__ void dmma_8x8x4(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

struct EpiParams {
  double alpha[2], beta[2];
  int beta_zero;
  int conja, conjb;
};

// grouped tile order: consecutive CTAs walk GROUP tile-rows first so that a wave of 148 CTAs shares ~16 A panels
// and ~10 B panels in L2.
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

template <bool CPLX>
__global__ void __launch_bounds__(THREADS, 1)
dmma_gemm_kernel(PanelSrc a, PanelSrc b, int64_t m, int64_t n, int64_t k, double* __restrict__ C, int64_t ldc,
                 EpiParams ep, int64_t tiles_m, int64_t tiles_n) {
  extern __shared__ __align__(16) double smem[];
  double* As = smem;
  double* Bs = smem + STAGES * PANEL;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp & 1, wn = warp >> 1;  // 2 x 4 warps, 64 x 32 warp tiles
  int64_t tm, tn;
  tile_of(blockIdx.x, tiles_m, tiles_n, tm, tn);
  constexpr int SC = CPLX ? 2 : 1;          // real rows/cols per scalar
  const int64_t m0 = tm * (BM / SC), n0 = tn * (BN / SC);  // tile origin in scalars

  double acc[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

  const int64_t nkt = (k + BK - 1) / BK;
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nkt) {
      load_panel<CPLX>(As + s * PANEL, a, m0, (int64_t)s * BK, k, tid);
      load_panel<CPLX>(Bs + s * PANEL, b, n0, (int64_t)s * BK, k, tid);
    }
    cp_async_commit();
  }

  const int fr = lane >> 2, fk = lane & 3;  // fragment row (0..7) and k (0..3) of this lane
  for (int64_t kt = 0; kt < nkt; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      const int64_t nxt = kt + STAGES - 1;
      if (nxt < nkt) {
        const int s = (int)(nxt % STAGES);
        load_panel<CPLX>(As + s * PANEL, a, m0, nxt * BK, k, tid);
        load_panel<CPLX>(Bs + s * PANEL, b, n0, nxt * BK, k, tid);
      }
      cp_async_commit();
    }
    const double* As_ = As + (kt % STAGES) * PANEL + wm * 64 + fr;
    const double* Bs_ = Bs + (kt % STAGES) * PANEL + wn * 32 + fr;
#pragma unroll
    for (int k4 = 0; k4 < BK / 4; ++k4) {
      double af[8], bf[4];
#pragma unroll
      for (int i = 0; i < 8; ++i) af[i] = As_[(k4 * 4 + fk) * LDS_ROW + i * 8];
#pragma unroll
      for (int j = 0; j < 4; ++j) bf[j] = Bs_[(k4 * 4 + fk) * LDS_ROW + j * 8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma_8x8x4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
  }
  cp_async_wait<0>();

  // ---- epilogue: registers -> global, alpha/beta fused ------------------------------------------------------
  // accumulator (i,j): real row wm*64 + i*8 + lane/4, real cols wn*32 + j*8 + 2*(lane%4) + {0,1}
  if constexpr (!CPLX) {
    const double alpha = ep.alpha[0], beta = ep.beta[0];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int64_t gj = n0 + wn * 32 + j * 8 + 2 * fk + c;
        if (gj >= n) continue;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int64_t gi = m0 + wm * 64 + i * 8 + fr;
          if (gi >= m) continue;
          double* pc = C + gi + gj * ldc;
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
    for (int j = 0; j < 4; ++j) {
      const int64_t gj = n0 + wn * 16 + j * 4 + fk;  // complex column
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int64_t gi = m0 + wm * 32 + i * 4 + (fr >> 1);  // complex row
        const double c0 = acc[i][j][0], c1 = acc[i][j][1];
        const double o1 = __shfl_xor_sync(0xffffffffu, c1, 4);
        // even lane: re = P + sS*S (P = own c0, S = partner c1); odd lane: im = sQ*Q + sR*R (Q = partner c1, R = own c0)
        const double mine = odd ? fma(sQ, o1, sR * c0) : fma(sS, o1, c0);
        const double other = __shfl_xor_sync(0xffffffffu, mine, 4);
        const double re = odd ? other : mine, im = odd ? mine : other;
        if (gi >= m || gj >= n) continue;
        double* pc = C + 2 * (gi + gj * ldc);
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

}  // namespace

bool dmma_supported(const GemmProblem& p) {
  if (p.type != TY_D && p.type != TY_Z) return false;
  if (((uintptr_t)p.A & 7) || ((uintptr_t)p.B & 7) || ((uintptr_t)p.C & 7)) return false;
  return p.m > 0 && p.n > 0 && p.k > 0;
}

int launch_dmma(const GemmProblem& p, cudaStream_t s) {
  const bool cplx = p.type == TY_Z;
  const int sc = cplx ? 2 : 1;
  PanelSrc a, b;
  a.base = (const double*)p.A; a.ld = p.lda; a.dim = p.m; a.dim_contig = (p.opa == OP_N);
  b.base = (const double*)p.B; b.ld = p.ldb; b.dim = p.n; b.dim_contig = (p.opb != OP_N);
  // 16-byte cp.async: real panels need a contiguous tile dimension, an even ld and a 16-byte aligned base;
  // complex scalars are 16 bytes themselves, so only the base alignment matters.
  a.vec16 = cplx ? aligned16(p.A) : (a.dim_contig && aligned16(p.A) && (p.lda % 2 == 0));
  b.vec16 = cplx ? aligned16(p.B) : (b.dim_contig && aligned16(p.B) && (p.ldb % 2 == 0));
  EpiParams ep;
  ep.alpha[0] = p.alpha[0]; ep.alpha[1] = p.alpha[1];
  ep.beta[0] = p.beta[0]; ep.beta[1] = p.beta[1];
  ep.beta_zero = (p.beta[0] == 0.0 && p.beta[1] == 0.0);
  ep.conja = (p.opa == OP_C); ep.conjb = (p.opb == OP_C);
  const int64_t tiles_m = (p.m * sc + BM - 1) / BM, tiles_n = (p.n * sc + BN - 1) / BN;
  const int64_t tiles = tiles_m * tiles_n;
  if (tiles > 0x7fffffffLL) return (int)cudaErrorInvalidConfiguration;
  static bool attr_done[2] = {false, false};
  if (cplx) {
    if (!attr_done[1]) {
      B200_CUDA_TRY(cudaFuncSetAttribute(dmma_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
      attr_done[1] = true;
    }
    note_variant("dmma_z_64x64x16");
    dmma_gemm_kernel<true><<<(unsigned)tiles, THREADS, SMEM_BYTES, s>>>(a, b, p.m, p.n, p.k, (double*)p.C, p.ldc, ep, tiles_m, tiles_n);
  } else {
    if (!attr_done[0]) {
      B200_CUDA_TRY(cudaFuncSetAttribute(dmma_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
      attr_done[0] = true;
    }
    note_variant("dmma_d_128x128x16");
    dmma_gemm_kernel<false><<<(unsigned)tiles, THREADS, SMEM_BYTES, s>>>(a, b, p.m, p.n, p.k, (double*)p.C, p.ldc, ep, tiles_m, tiles_n);
  }
  count_launch();
  return (int)cudaGetLastError();
}

}  // namespace b200
