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
// Tile configuration: 128x64x16 CTA tile, 8 warps of 32x32, 2 CTAs per SM (chosen from the sweep in
// profiles/variant_sweep_r01.md over 128x128/1-CTA, 128x64/2-CTA, 16-warp and BK=32 variants).  The loader mode of
// each operand (16-byte / 8-byte / transposed) is a template parameter so that no mode state occupies registers.
#include <cstdlib>
#include <type_traits>

#include "common.cuh"

namespace b200 {
namespace {

template <int BM_, int BN_, int WM_, int WN_, int MINB_, int BK_ = 16, int STAGES_ = 4, bool MBAR_ = false>
struct Cfg {
  static constexpr int BM = BM_, BN = BN_, WM = WM_, WN = WN_, MINB = MINB_, BK = BK_, STAGES = STAGES_;
  // MBAR: per-stage full/empty mbarriers instead of one __syncthreads per k-tile.  Copies run two tiles ahead in a
  // four-stage ring, so a warp only waits for warps that are more than a whole tile behind it.
  static constexpr bool MBAR = MBAR_;
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

// Loader modes (compile time, so that no mode state lives in registers):
//   LD_DIM16  memory contiguous along the tile dimension, 16-byte cp.async (real: 2 rows per copy; complex: 1 scalar)
//   LD_DIM8   memory contiguous along the tile dimension, 8-byte cp.async (any alignment / odd ld; real only)
//   LD_K      memory contiguous along k (transposed operand): 8-byte copies (real) or 16-byte scalars (complex)
enum { LD_DIM16 = 0, LD_DIM8 = 1, LD_K = 2 };

// Per-thread loader for one operand panel of PR real rows: copy e (0 <= e < E) moves BYTES bytes from
// p + e*estep (+ tile offset) to smem offset soff + e*SSTEP and belongs to k index kk0 + e*KKSTEP of the tile.
template <int PR, int THREADS, bool CPLX, int BK, int MODE, int LDS_ROW>
struct Loader {
  static constexpr int GRAN = (CPLX || MODE == LD_DIM16) ? 2 : 1;   // doubles per copy
  static constexpr int RS = CPLX ? 2 : 1;                             // doubles per scalar
  static constexpr int GD = PR / GRAN;                                // granules along the tile dimension
  static constexpr int BYTES = 8 * GRAN;
  static constexpr int E = GD * BK / THREADS;
  static constexpr bool DIMC = (MODE != LD_K);
  static constexpr int KKSTEP = DIMC ? THREADS / GD : 0;
  static constexpr int RSTEP = THREADS / BK;                          // LD_K: scalars between consecutive copies
  static constexpr int SSTEP = DIMC ? KKSTEP * LDS_ROW : RSTEP * RS;
  static_assert(E >= 1 && E <= 16, "copies per thread");
  static_assert(DIMC ? (THREADS % GD == 0) : (THREADS % BK == 0), "thread map");
  static_assert(!(CPLX && MODE == LD_DIM8), "complex scalars move as 16-byte copies");

  const double* p;   // source of copy 0 of k-tile 0
  uint32_t soff;     // smem offset (doubles) of copy 0
  uint32_t vmask;    // LD_K: bit e = row/column of copy e inside the matrix; DIM modes: bytes valid (0, 8, 16)
  int kk0;           // k index of copy 0 inside the tile

  __device__ __forceinline__ void init(const PanelSrc& s, int64_t dim0, int tid) {
    if constexpr (DIMC) {
      const int rg = tid % GD;
      kk0 = tid / GD;
      const int64_t gd = dim0 + (int64_t)rg * GRAN / RS;   // first scalar of the granule
      p = s.base + RS * gd + (int64_t)kk0 * s.ld * RS;
      soff = (uint32_t)(kk0 * LDS_ROW + rg * GRAN);
      vmask = 0;
      if (gd < s.dim) vmask = (!CPLX && GRAN == 2 && gd + 1 >= s.dim) ? 8u : (uint32_t)BYTES;
    } else {
      kk0 = tid % BK;
      const int r0 = tid / BK;
      p = s.base + RS * ((int64_t)kk0 + (dim0 + r0) * s.ld);
      soff = (uint32_t)(kk0 * LDS_ROW + r0 * RS);
      vmask = 0;
#pragma unroll
      for (int e = 0; e < E; ++e)
        if (dim0 + r0 + (int64_t)e * RSTEP < s.dim) vmask |= 1u << e;
    }
  }
  __device__ __forceinline__ int64_t estep(const PanelSrc& s) const {   // doubles between consecutive copies
    return DIMC ? (int64_t)KKSTEP * s.ld * RS : (int64_t)RSTEP * s.ld * RS;
  }
  __device__ __forceinline__ int64_t kstep(const PanelSrc& s) const {   // doubles between consecutive k-tiles
    return DIMC ? (int64_t)BK * s.ld * RS : (int64_t)BK * RS;
  }
  // Copy e of an interior k-tile (all BK k values exist): q = source of THIS copy (the caller advances it by estep).
  // Rows / columns outside the matrix are simply not loaded: whatever the slot holds only ever reaches rows / columns
  // of C that the epilogue masks (every output element is its own dot product).
  __device__ __forceinline__ void copy_fast(double* S, const double* q, int e) const {
    double* dst = S + soff + e * SSTEP;
    if constexpr (DIMC) {
      if (vmask == (uint32_t)BYTES) cp_async_full<BYTES>(dst, q);
      else if (vmask) cp_async_zfill<BYTES>(dst, q, (int)vmask);   // odd last row of a 16-byte real granule
    } else {
      if ((vmask >> e) & 1u) cp_async_full<BYTES>(dst, q);
    }
  }
  // Copy e of the ragged last k-tile: k values beyond the matrix must read as zero (they reach every output)
  __device__ __forceinline__ void copy_ragged(double* S, const double* q, const double* fallback, int e, int krem) const {
    int nbytes;
    if constexpr (DIMC) nbytes = (kk0 + e * KKSTEP < krem) ? (int)vmask : 0;
    else nbytes = (((vmask >> e) & 1u) && kk0 < krem) ? BYTES : 0;
    cp_async_zfill<BYTES>(S + soff + e * SSTEP, nbytes ? q : fallback, nbytes);
  }
};

__device__ __forceinline__ void dmma_8x8x4(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  const uint32_t a = smem_u32(bar);
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done) : "r"(a), "r"(parity) : "memory");
  } while (!done);
}
// one non-blocking probe: 1 if the phase with the given parity has completed
__device__ __forceinline__ uint32_t mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
               : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return done;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// arrive on `bar` once all cp.async copies issued so far by this thread have landed (no pending-count increment)
__device__ __forceinline__ void cp_async_mbar_arrive(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct EpiParams {
  double alpha[2], beta[2];
  int beta_zero;
  int conja, conjb;
  int uplo, herm;   // triangular mask of ?syrk_/?herk_ (common.cuh)
  int diag;   // B200BLAS_DMMA_DIAG (measurement only, results are garbage): 1 = no copies, no waits; 2 = copies, no waits
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

// inverse of tile_of: position of tile (tm, tn) in the launch order
__device__ __forceinline__ int64_t pid_of(int64_t tm, int64_t tn, int64_t tiles_m, int64_t tiles_n) {
  const int64_t g = tm / GROUP, first = g * GROUP;
  const int64_t gsz = (tiles_m - first < GROUP) ? (tiles_m - first) : GROUP;
  return g * GROUP * tiles_n + tn * gsz + (tm - first);
}
// Wave-quantisation tail: the main launch covers the first `covered` tiles of the big-tile grid (a whole number of waves),
// a second launch with the small tile covers the rest and skips everything the main launch owns.
struct TailSkip {
  int64_t covered = 0;                   // 0: ordinary launch
  int64_t big_tiles_m = 0, big_tiles_n = 0;
  int ratio_m = 1, ratio_n = 1;          // small tiles per big tile along m / n
};

// TRI: the triangular mask of ?syrk_/?herk_/?syr2k_/?her2k_ is compiled in.  The plain product is a separate
// instantiation without a single mask instruction: with the mask tests in the epilogue the 16384^3 dgemm ran 1.8 %
// slower (A/B on the same B200: 257.0 -> 261.8 ms), although the mask is never active there.
template <typename C, bool CPLX, int AMODE, int BMODE, bool TRI>
__global__ void __launch_bounds__(C::THREADS, C::MINB)
dmma_gemm_kernel(PanelSrc a, PanelSrc b, int64_t m, int64_t n, int64_t k, double* __restrict__ Cmat, int64_t ldc,
                 EpiParams ep, int64_t tiles_m, int64_t tiles_n, TailSkip skip) {
  extern __shared__ __align__(16) double smem[];
  double* As = smem;
  double* Bs = smem + C::STAGES * C::PANEL_A;
  constexpr int MI = C::MI, NJ = C::NJ;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp % C::WARPS_M, wn = warp / C::WARPS_M;
  int64_t tm, tn;
  tile_of(blockIdx.x, tiles_m, tiles_n, tm, tn);
  // tail launch (launch_split_tail): tiles already covered by the first `covered` big tiles of the main launch are not ours
  if (skip.covered > 0 && pid_of(tm / skip.ratio_m, tn / skip.ratio_n, skip.big_tiles_m, skip.big_tiles_n) < skip.covered) return;
  constexpr int SC = CPLX ? 2 : 1;  // real rows/cols per scalar
  const int64_t m0 = tm * (C::BM / SC), n0 = tn * (C::BN / SC);  // tile origin in scalars
  if constexpr (TRI) {
    if (tile_outside(ep.uplo, m0, m0 + C::BM / SC, n0, n0 + C::BN / SC)) return;   // rank-k update: other triangle
  }
  auto in_tri = [&](int64_t i, int64_t j) -> bool {
    if constexpr (TRI) return in_triangle(ep.uplo, i, j); else return true;
  };

  // short k (rank-k updates: config C5, the trailing updates of ?potrf_ / ?getrf_): the read-modify-write of the C tile in the
  // epilogue is a full HBM round trip that nothing overlaps once the k loop is over -- request the tile's lines into L2 now
  if (!ep.beta_zero && k <= 1024) {
    constexpr int LINES_PER_COL = C::BM / 16, NCOLS = C::BN / SC;   // BM * 8 bytes per column = BM / 16 lines of 128 bytes
    for (int line = tid; line < LINES_PER_COL * NCOLS; line += C::THREADS) {
      const int64_t gi = m0 + (int64_t)(line % LINES_PER_COL) * (16 / SC), gj = n0 + line / LINES_PER_COL;
      if (gi < m && gj < n) asm volatile("prefetch.global.L2 [%0];" ::"l"(Cmat + SC * (gi + gj * ldc)));
    }
  }
  constexpr int BK = C::BK, STAGES = C::STAGES;
  using LA = Loader<C::BM, C::THREADS, CPLX, BK, AMODE, C::LDA_S>;
  using LB = Loader<C::BN, C::THREADS, CPLX, BK, BMODE, C::LDB_S>;
  LA la;
  LB lb;
  la.init(a, m0, tid);
  lb.init(b, n0, tid);

  double acc[MI][NJ][2];
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NJ; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

  const int64_t nkt = (k + BK - 1) / BK;
  // running source pointers: pa_next / pb_next = copy 0 of the next tile to be loaded; the per-copy and per-tile strides
  // are loop invariant (no 64-bit multiplies in the loop)
  const int64_t ea = la.estep(a), eb = lb.estep(b), ka = la.kstep(a), kb = lb.kstep(b);
  const double* pa_next = la.p;
  const double* pb_next = lb.p;
  auto load_all = [&](int64_t kt) {
    double* sa = As + (kt % STAGES) * C::PANEL_A;
    double* sb = Bs + (kt % STAGES) * C::PANEL_B;
    const int64_t rem = k - kt * BK;
    const double* qa = pa_next;
    const double* qb = pb_next;
    if (rem >= BK) {
#pragma unroll
      for (int e = 0; e < LA::E; ++e) { la.copy_fast(sa, qa, e); qa += ea; }
#pragma unroll
      for (int e = 0; e < LB::E; ++e) { lb.copy_fast(sb, qb, e); qb += eb; }
    } else {
#pragma unroll
      for (int e = 0; e < LA::E; ++e) { la.copy_ragged(sa, qa, a.base, e, (int)rem); qa += ea; }
#pragma unroll
      for (int e = 0; e < LB::E; ++e) { lb.copy_ragged(sb, qb, b.base, e, (int)rem); qb += eb; }
    }
    pa_next += ka;
    pb_next += kb;
  };
  constexpr int LOOK = C::MBAR ? STAGES - 2 : STAGES - 1;   // tiles in flight ahead of the one being multiplied
  __shared__ uint64_t full_bar[C::MBAR ? STAGES : 1], empty_bar[C::MBAR ? STAGES : 1];
  if constexpr (C::MBAR) {
    if (tid == 0) {
#pragma unroll
      for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], C::THREADS); mbar_init(&empty_bar[s], C::THREADS / 32); }
    }
    __syncthreads();
  }
#pragma unroll
  for (int s = 0; s < LOOK; ++s) {
    if (s < nkt) {
      load_all(s);
      if constexpr (C::MBAR) cp_async_mbar_arrive(&full_bar[s]);
    }
    if constexpr (!C::MBAR) cp_async_commit();
  }

  const int fr = lane >> 2, fk = lane & 3;  // fragment row (0..7) and k (0..3) of this lane
  // The barrier probes for tile kt+1 are issued BEFORE the last 16 DMMAs of tile kt and only checked after them, so
  // the ~100-cycle mbarrier round trip is hidden under MMA issue instead of stalling every warp at each tile boundary
  // (measured: main loop without copies/waits 35.9 TFLOP/s, with blocking waits at the boundary 32.6).
  uint32_t pre_full = 0, pre_empty = 0;
  for (int64_t kt = 0; kt < nkt; ++kt) {
    const int64_t nxt = kt + LOOK;
    const bool do_load = nxt < nkt && ep.diag != 1;
    const int st = (int)(kt % STAGES), sn = (int)(nxt % STAGES);
    if constexpr (C::MBAR) {
      // the ring slot of tile nxt was last read by tile nxt - STAGES: every warp must have released it
      if (!ep.diag) {
        if (do_load && nxt >= STAGES && !pre_empty) mbar_wait(&empty_bar[sn], (uint32_t)((nxt / STAGES - 1) & 1));
        if (!pre_full) mbar_wait(&full_bar[st], (uint32_t)((kt / STAGES) & 1));
      }
    } else {
      cp_async_wait<STAGES - 2>();
      __syncthreads();
    }
    double* sa_n = As + sn * C::PANEL_A;
    double* sb_n = Bs + sn * C::PANEL_B;
    const int64_t rem_n = k - nxt * BK;
    const bool fast = rem_n >= BK;   // interior tile: unpredicated-size copies
    const double* qa = pa_next;
    const double* qb = pb_next;
    const double* As_ = As + st * C::PANEL_A + wm * C::WM + fr;
    const double* Bs_ = Bs + st * C::PANEL_B + wn * C::WN + fr;
    // MODE 0: no copies ride along (tail of the k loop / diagnostic); 1: interior tile (plain copies); 2: ragged last tile
    auto tile_body = [&](auto mode_tag) {
      constexpr int MODE = decltype(mode_tag)::value;
#pragma unroll
      for (int k4 = 0; k4 < BK / 4; ++k4) {
        double af[MI], bf[NJ];
#pragma unroll
        for (int i = 0; i < MI; ++i) af[i] = As_[(k4 * 4 + fk) * C::LDA_S + i * 8];
#pragma unroll
        for (int j = 0; j < NJ; ++j) bf[j] = Bs_[(k4 * 4 + fk) * C::LDB_S + j * 8];
        // a quarter of the next tile's copies rides along with each k4 step
        if constexpr (MODE != 0) {
#pragma unroll
          // copies are issued in increasing e (the source pointers are running sums): k4 step j takes the j-th chunk
          constexpr int PER_A = (LA::E + BK / 4 - 1) / (BK / 4), PER_B = (LB::E + BK / 4 - 1) / (BK / 4);
#pragma unroll
          for (int e = 0; e < LA::E; ++e) {
            if (e / PER_A == k4) {
              if constexpr (MODE == 1) la.copy_fast(sa_n, qa, e); else la.copy_ragged(sa_n, qa, a.base, e, (int)rem_n);
              qa += ea;
            }
          }
#pragma unroll
          for (int e = 0; e < LB::E; ++e) {
            if (e / PER_B == k4) {
              if constexpr (MODE == 1) lb.copy_fast(sb_n, qb, e); else lb.copy_ragged(sb_n, qb, b.base, e, (int)rem_n);
              qb += eb;
            }
          }
        }
        if constexpr (C::MBAR) {
          if (k4 == BK / 4 - 1) {
            // this thread's copies of tile nxt are all issued: publish them, then probe the barriers of tile kt+1
            if (MODE != 0) cp_async_mbar_arrive(&full_bar[sn]);
            pre_full = pre_empty = 0;
            if (!ep.diag && kt + 1 < nkt) {
              pre_full = mbar_try(&full_bar[(kt + 1) % STAGES], (uint32_t)(((kt + 1) / STAGES) & 1));
              const int64_t n2 = nxt + 1;
              if (n2 < nkt && n2 >= STAGES) pre_empty = mbar_try(&empty_bar[n2 % STAGES], (uint32_t)((n2 / STAGES - 1) & 1));
            }
          }
        }
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
          for (int j = 0; j < NJ; ++j) dmma_8x8x4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
      }
    };
    if (!do_load) tile_body(std::integral_constant<int, 0>{});
    else if (fast) tile_body(std::integral_constant<int, 1>{});
    else tile_body(std::integral_constant<int, 2>{});
    if (do_load) { pa_next += ka; pb_next += kb; }
    if constexpr (C::MBAR) {
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[st]);   // this warp has read everything it needs from slot st
    } else {
      cp_async_commit();
    }
  }
  cp_async_wait<0>();

  // ---- epilogue: registers -> global, alpha/beta fused ------------------------------------------------------
  // accumulator (i,j): real row wm*WM + i*8 + lane/4, real cols wn*WN + j*8 + 2*(lane%4) + {0,1}
  if constexpr (!CPLX) {
    const double alpha = ep.alpha[0], beta = ep.beta[0];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      // loads of this column pair first, stores afterwards: a load placed after a store through the same pointer
      // cannot be hoisted by the compiler, which would serialise every read-modify-write round trip
      double old[2][MI];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int64_t gj = n0 + wn * C::WN + j * 8 + 2 * fk + c;
#pragma unroll
        for (int i = 0; i < MI; ++i) {
          const int64_t gi = m0 + wm * C::WM + i * 8 + fr;
          old[c][i] = (!ep.beta_zero && gi < m && gj < n && in_tri(gi, gj)) ? Cmat[gi + gj * ldc] : 0.0;
        }
      }
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int64_t gj = n0 + wn * C::WN + j * 8 + 2 * fk + c;
#pragma unroll
        for (int i = 0; i < MI; ++i) {
          const int64_t gi = m0 + wm * C::WM + i * 8 + fr;
          if (gi < m && gj < n && in_tri(gi, gj)) Cmat[gi + gj * ldc] = fma(beta, old[c][i], alpha * acc[i][j][c]);
        }
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
      double outv[MI];
#pragma unroll
      for (int i = 0; i < MI; ++i) {
        const double c0 = acc[i][j][0], c1 = acc[i][j][1];
        const double o1 = __shfl_xor_sync(0xffffffffu, c1, 4);
        // even lane: re = P + sS*S (P = own c0, S = partner c1); odd lane: im = sQ*Q + sR*R (Q = partner c1, R = own c0)
        const double mine = odd ? fma(sQ, o1, sR * c0) : fma(sS, o1, c0);
        const double other = __shfl_xor_sync(0xffffffffu, mine, 4);
        const double re = odd ? other : mine, im = odd ? mine : other;
        // out = alpha*(re,im) + beta*Cold ; even lane stores the real part, odd lane the imaginary part
        outv[i] = odd ? fma(ep.alpha[0], im, ep.alpha[1] * re) : fma(ep.alpha[0], re, -ep.alpha[1] * im);
      }
      // loads of the whole column first, stores afterwards (no load is stuck behind a store to the same array)
      if (!ep.beta_zero) {
        double cr[MI], ci[MI];
#pragma unroll
        for (int i = 0; i < MI; ++i) {
          const int64_t gi = m0 + wm * (C::WM / 2) + i * 4 + (fr >> 1);
          const bool ok = gi < m && gj < n && in_tri(gi, gj);
          cr[i] = ok ? Cmat[2 * (gi + gj * ldc)] : 0.0;
          ci[i] = ok ? Cmat[2 * (gi + gj * ldc) + 1] : 0.0;
        }
#pragma unroll
        for (int i = 0; i < MI; ++i)
          outv[i] += odd ? fma(ep.beta[0], ci[i], ep.beta[1] * cr[i]) : fma(ep.beta[0], cr[i], -ep.beta[1] * ci[i]);
      }
#pragma unroll
      for (int i = 0; i < MI; ++i) {
        const int64_t gi = m0 + wm * (C::WM / 2) + i * 4 + (fr >> 1);
        if (gi < m && gj < n && in_tri(gi, gj))
          Cmat[2 * (gi + gj * ldc) + (odd ? 1 : 0)] = (TRI && odd && ep.herm && gi == gj) ? 0.0 : outv[i];
      }
    }
  }
}

bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

// 128x64x16 tile, 8 warps of 32x32, 4-stage ring, 2 CTAs / SM: the two co-resident CTAs cover each other's barrier,
// prologue and epilogue bubbles (profiles/variant_sweep_r01.md: wins or ties against 128x128 / 1 CTA per SM).
using CfgB = Cfg<128, 64, 32, 32, 2, 16, 4, false>;
using CfgM = Cfg<128, 64, 32, 32, 2, 16, 4, true>;   // same tile, mbarrier pipeline (B200BLAS_DMMA_SYNC=mbar)
// 64x32x16 tile, 8 warps of 16x16, 3 CTAs / SM: for products whose 128x64 tiling fills less than a third of the 296
// CTA slots (the mid-size updates inside ?trsm_ / ?potrf_ / ?getrf_): a CTA's k-loop is bound by its own DMMA issue rate
// (64 DMMA per warp and k-tile = 2048 cycles with two warps per sub-partition), so a quarter-size tile on four times as
// many SMs finishes the same product up to 4x sooner.
using CfgS = Cfg<64, 32, 16, 16, 3, 16, 4, true>;

// grid_limit > 0: only the first grid_limit tiles of the launch order (the main part of a split launch)
template <typename C, bool CPLX, int AMODE, int BMODE>
int launch_cfg(const GemmProblem& p, cudaStream_t s, const PanelSrc& a, const PanelSrc& b, const EpiParams& ep,
               int64_t grid_limit = 0, const TailSkip& skip = TailSkip()) {
  const int sc = CPLX ? 2 : 1;
  const int64_t tiles_m = (p.m * sc + C::BM - 1) / C::BM, tiles_n = (p.n * sc + C::BN - 1) / C::BN;
  int64_t tiles = tiles_m * tiles_n;
  if (tiles > 0x7fffffffLL) return (int)cudaErrorInvalidConfiguration;
  if (grid_limit > 0 && grid_limit < tiles) tiles = grid_limit;
  if (p.uplo != UPLO_FULL) {
    B200_SET_MAX_DYN_SMEM_ONCE((dmma_gemm_kernel<C, CPLX, AMODE, BMODE, true>), C::SMEM_BYTES);
    dmma_gemm_kernel<C, CPLX, AMODE, BMODE, true><<<(unsigned)tiles, C::THREADS, C::SMEM_BYTES, s>>>(
        a, b, p.m, p.n, p.k, (double*)p.C, p.ldc, ep, tiles_m, tiles_n, skip);
  } else {
    B200_SET_MAX_DYN_SMEM_ONCE((dmma_gemm_kernel<C, CPLX, AMODE, BMODE, false>), C::SMEM_BYTES);
    dmma_gemm_kernel<C, CPLX, AMODE, BMODE, false><<<(unsigned)tiles, C::THREADS, C::SMEM_BYTES, s>>>(
        a, b, p.m, p.n, p.k, (double*)p.C, p.ldc, ep, tiles_m, tiles_n, skip);
  }
  count_launch();
  return (int)cudaGetLastError();
}

template <typename CF, bool CPLX, int AMODE>
int launch_b(int bmode, const GemmProblem& p, cudaStream_t s, const PanelSrc& a, const PanelSrc& b, const EpiParams& ep,
             int64_t grid_limit, const TailSkip& skip) {
  if constexpr (CPLX) {
    return bmode == LD_K ? launch_cfg<CF, true, AMODE, LD_K>(p, s, a, b, ep, grid_limit, skip)
                         : launch_cfg<CF, true, AMODE, LD_DIM16>(p, s, a, b, ep, grid_limit, skip);
  } else {
    switch (bmode) {
      case LD_DIM16: return launch_cfg<CF, false, AMODE, LD_DIM16>(p, s, a, b, ep, grid_limit, skip);
      case LD_DIM8: return launch_cfg<CF, false, AMODE, LD_DIM8>(p, s, a, b, ep, grid_limit, skip);
      default: return launch_cfg<CF, false, AMODE, LD_K>(p, s, a, b, ep, grid_limit, skip);
    }
  }
}

template <typename CF>
int launch_modes(bool cplx, int amode, int bmode, const GemmProblem& p, cudaStream_t s, const PanelSrc& a, const PanelSrc& b,
                 const EpiParams& ep, int64_t grid_limit = 0, const TailSkip& skip = TailSkip()) {
  if (cplx) return amode == LD_K ? launch_b<CF, true, LD_K>(bmode, p, s, a, b, ep, grid_limit, skip)
                                 : launch_b<CF, true, LD_DIM16>(bmode, p, s, a, b, ep, grid_limit, skip);
  switch (amode) {
    case LD_DIM16: return launch_b<CF, false, LD_DIM16>(bmode, p, s, a, b, ep, grid_limit, skip);
    case LD_DIM8: return launch_b<CF, false, LD_DIM8>(bmode, p, s, a, b, ep, grid_limit, skip);
    default: return launch_b<CF, false, LD_K>(bmode, p, s, a, b, ep, grid_limit, skip);
  }
}

bool use_mbar() {
  // default: mbarrier pipeline (<= 1.5 % faster than __syncthreads in profiles/variant_sweep_r01.md); "B200BLAS_DMMA_SYNC=bar" selects the other
  static const bool v = [] { const char* e = getenv("B200BLAS_DMMA_SYNC"); return !(e && e[0] == 'b'); }();
  return v;
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
  ep.uplo = p.uplo; ep.herm = p.herm;
  static const int diag_env = [] { const char* e = getenv("B200BLAS_DMMA_DIAG"); return e ? atoi(e) : 0; }();
  ep.diag = diag_env;
  // loader modes: 16-byte copies need the tile dimension contiguous, a 16-byte aligned base and (real) an even ld
  const int amode = !a.dim_contig ? LD_K : (a.vec16 ? LD_DIM16 : LD_DIM8);
  const int bmode = !b.dim_contig ? LD_K : (b.vec16 ? LD_DIM16 : LD_DIM8);
  // k-slicing: one wave of 296 CTAs streams (16*128 + 18.5*64) panel rows x k x 8 bytes; beyond ~4096 k that window no
  // longer fits the 126 MB L2 and panel reuse between CTAs depends on them staying in lockstep (measured at 16384^3:
  // L2 hit rate 83 % -> 53 % and 82 -> 338 GB of DRAM reads once the CTAs drift).  Large-k products therefore run as
  // consecutive k-slices of 2048 (beta = 1 after the first), which is also the reference's kc blocking
  // (GeneralMatrixMatrix.h:172-174): every slice ends in one rounding into C.  Measured at 16384^3
  // (profiles/variant_sweep_r01.md): unsliced 265.5 ms / 340 GB of DRAM traffic, 4096-slices 266.8 ms / 102 GB,
  // 2048-slices 268.1 ms / 57.6 GB (L2 hit rate 92 %).
  static const int64_t kslice_env = [] { const char* e = getenv("B200BLAS_DMMA_KSLICE"); return e ? (int64_t)atoll(e) : (int64_t)-1; }();
  const int64_t sc = cplx ? 2 : 1;
  int64_t kslice = kslice_env >= 0 ? kslice_env : 2048;
  const bool big_panels = (double)(p.m + p.n) * sc > 6000.0;   // small m+n: the panels of all tiles fit L2 anyway
  if (kslice <= 0 || !big_panels || p.k < 2 * kslice) kslice = p.k;
  const int64_t nslices = (p.k + kslice / 2) / kslice > 0 ? (p.k + kslice / 2) / kslice : 1;   // nearest count
  const int64_t slice_len = ((p.k + nslices - 1) / nslices + 15) / 16 * 16;                   // equal slices, BK-aligned
  for (int64_t k0 = 0; k0 < p.k; k0 += slice_len) {
    GemmProblem q = p;
    q.k = p.k - k0 < slice_len ? p.k - k0 : slice_len;
    PanelSrc a2 = a, b2 = b;
    a2.base = a.base + sc * (a.dim_contig ? k0 * p.lda : k0);
    b2.base = b.base + sc * (b.dim_contig ? k0 * p.ldb : k0);
    EpiParams e2 = ep;
    if (k0 > 0) { e2.beta[0] = 1.0; e2.beta[1] = 0.0; e2.beta_zero = 0; }
    int rc;
    static const int tile_env = [] { const char* e = getenv("B200BLAS_DMMA_TILE"); return !e ? 0 : (e[0] == 's' ? 1 : 2); }();
    const int64_t big_tiles = ((q.m * sc + CfgM::BM - 1) / CfgM::BM) * ((q.n * sc + CfgM::BN - 1) / CfgM::BN);
    // up to 222 big tiles (three quarters of one wave of 296) the quarter-size tile finishes sooner: 4r small tiles are at
    // most two waves of ~3/8 the duration each (measured round 1: 1024^3 23.6 vs 22.9 TF, 1536^3 = 288 tiles 29.4 vs 30.3)
    if (tile_env == 1 || (tile_env == 0 && big_tiles <= 222)) {
      note_variant(cplx ? "dmma_z_32x16x16_w8x8_3cta_mbar" : "dmma_d_64x32x16_w16x16_3cta_mbar");
      rc = launch_modes<CfgS>(cplx, amode, bmode, q, s, a2, b2, e2);
    } else if (use_mbar()) {
      // Wave quantisation: the big tile runs 2 CTAs per SM = `slots` tiles per wave.  When the last wave would be at most
      // three quarters full, the main launch stops at a whole number of waves and the rest runs on the quarter-size tile
      // (3 CTAs per SM): r big tiles = 4r small ones = ceil(4r / (3 * SMs)) waves of ~3/8 the duration.  2048^3: 512 tiles
      // = 1.73 waves -> 1 + 2 * 0.375 instead of 2 (B200BLAS_DMMA_TAIL=0 disables).
      static const bool tail_on = [] { const char* e = getenv("B200BLAS_DMMA_TAIL"); return !(e && e[0] == '0'); }();
      static const int sms = [] { int d = 0, n = 148; if (cudaGetDevice(&d) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d); return n; }();
      const int64_t slots = 2 * (int64_t)sms, waves = big_tiles / slots, rest = big_tiles - waves * slots;
      const int64_t small_waves = (4 * rest + 3 * sms - 1) / (3 * sms);
      if (tail_on && q.uplo == UPLO_FULL && rest > 0 && waves >= 1 && waves <= 24 && 3 * small_waves < 8) {
        note_variant(cplx ? "dmma_z_64x32x16_w16x16_2cta_mbar+tail" : "dmma_d_128x64x16_w32x32_2cta_mbar+tail");
        rc = launch_modes<CfgM>(cplx, amode, bmode, q, s, a2, b2, e2, waves * slots);
        if (!rc) {
          TailSkip sk;
          sk.covered = waves * slots;
          sk.big_tiles_m = (q.m * sc + CfgM::BM - 1) / CfgM::BM; sk.big_tiles_n = (q.n * sc + CfgM::BN - 1) / CfgM::BN;
          sk.ratio_m = CfgM::BM / CfgS::BM; sk.ratio_n = CfgM::BN / CfgS::BN;
          rc = launch_modes<CfgS>(cplx, amode, bmode, q, s, a2, b2, e2, 0, sk);
        }
      } else {
        note_variant(cplx ? "dmma_z_64x32x16_w16x16_2cta_mbar" : "dmma_d_128x64x16_w32x32_2cta_mbar");
        rc = launch_modes<CfgM>(cplx, amode, bmode, q, s, a2, b2, e2);
      }
    } else {
      note_variant(cplx ? "dmma_z_64x32x16_w16x16_2cta" : "dmma_d_128x64x16_w32x32_2cta");
      rc = launch_modes<CfgB>(cplx, amode, bmode, q, s, a2, b2, e2);
    }
    if (rc) return rc;
  }
  return 0;
}

}  // namespace b200
