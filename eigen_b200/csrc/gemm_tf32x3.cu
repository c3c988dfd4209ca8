// eigen_b200/csrc/gemm_tf32x3.cu -- float GEMM on the 5th-gen tensor cores with a 3xTF32 split (sm_100a).
//
// Replaces the reference's float path gemm_pack_lhs/rhs + gebp_kernel
// (Eigen/src/Core/products/GeneralBlockPanelKernel.h:858-2105) for large sgemm products:
//   * pack (the analogue of gemm_pack_lhs / gemm_pack_rhs, :1688-2105): one pass over op(A) (m x k) and op(B)^T
//     (n x k) writes K-major panels split into hi = tf32(x) and lo = tf32(x - hi); the pass also canonicalises
//     N/T/C and any lda/ldb, so the product kernel only ever sees 128-byte aligned K-major rows (TMA-legal);
//   * product (the analogue of gebp_kernel, :858-1669): persistent, warp-specialised kernel -- warp 0 issues TMA
//     (cp.async.bulk.tensor, SWIZZLE_128B) into a 2-stage shared-memory ring, one elected thread of warp 1 issues
//     tcgen05.mma.kind::tf32 (M=128, N=256, K=8) three times per k-step -- lo*hi + hi*lo first, hi*hi last -- into
//     a double-buffered 128x256 fp32 accumulator in TMEM, warps 2-5 drain TMEM with tcgen05.ld and apply
//     C = alpha*acc + beta*C with coalesced column-major stores (the reference does beta in a separate pass,
//     blas/level3_impl.h:62-66).  The dropped lo*lo term is O(2^-22 |a||b|) per product, below fp32 rounding.
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"

namespace b200 {
namespace {

constexpr int TM = 128, TN = 256, TK = 32, UK = 8, NSTAGE = 2;
constexpr int A_PLANE = TM * TK * 4, B_PLANE = TN * TK * 4;
constexpr int STAGE_BYTES = 2 * A_PLANE + 2 * B_PLANE;  // 98304
constexpr int SMEM_BYTES = NSTAGE * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr int THREADS = 192;     // complex / 256-tile kernels: TMA warp, MMA warp, 4 drain warps
// real kernel: warpgroup 0 = TMA warp + MMA warp (+ 2 idle warps, setmaxnreg.dec 40); warpgroups 1-2 = 8 drain warps,
// two per TMEM lane quadrant (setmaxnreg.inc 232: 128 fp32 accumulators per thread stay in registers)
constexpr int THREADS_R = 384;
constexpr int TMEM_COLS = 512;

// ---------------- PTX wrappers ------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// K-major operand, 128-byte rows, SWIZZLE_128B: 8-row groups are 1024 bytes apart (SBO); LBO unused
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;              // leading byte offset (ignored for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;    // stride byte offset
  d |= (uint64_t)1 << 46;              // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;              // SWIZZLE_128B
  return d;
}
// instruction descriptor, kind::tf32: D = f32, A = B = tf32, both K-major
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// hi = tf32(x) (round to nearest), lo = tf32(x - hi).  Non-finite values must behave as in the reference's fp32 FMA path:
// x = +-Inf would give lo = Inf - Inf = NaN, and a finite |x| just below FLT_MAX rounds UP to Inf under cvt.rna -- the
// first keeps lo = 0 (Inf propagates through hi alone), the second truncates instead of rounding so that hi stays finite.
__device__ __forceinline__ void tf32_split(float x, float& hi, float& lo) {
  uint32_t h, l;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
  if ((h & 0x7f800000u) == 0x7f800000u) {
    const uint32_t xb = __float_as_uint(x);
    if ((xb & 0x7f800000u) != 0x7f800000u) h = xb & 0xffffe000u;
  }
  float res = x - __uint_as_float(h);
  if ((h & 0x7f800000u) == 0x7f800000u) res = 0.f;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(res));
  hi = __uint_as_float(h);
  lo = __uint_as_float(l);
}

// ---------------- pack: split op(X) into K-major tf32 hi / lo panels -----------------------------------------
// element (r, kk) of the logical R x K operand lives at src[r*sr + kk*sk]; exactly one of sr, sk is 1.
__global__ void __launch_bounds__(256)
tf32_split_pack_kernel(const float* __restrict__ src, int64_t sr, int64_t sk, int64_t R, int64_t K,
                       float* __restrict__ hi, float* __restrict__ lo, int64_t Kp) {
  __shared__ float t[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t r0 = (int64_t)blockIdx.y * 32, k0 = (int64_t)blockIdx.x * 32;
  if (sk == 1) {
#pragma unroll
    for (int j = ty; j < 32; j += 8) {
      const int64_t r = r0 + j, kk = k0 + tx;
      t[j][tx] = (r < R && kk < K) ? src[r * sr + kk] : 0.f;
    }
  } else {
#pragma unroll
    for (int j = ty; j < 32; j += 8) {
      const int64_t kk = k0 + j, r = r0 + tx;
      t[tx][j] = (r < R && kk < K) ? src[r + kk * sk] : 0.f;
    }
  }
  __syncthreads();
#pragma unroll
  for (int j = ty; j < 32; j += 8) {
    const int64_t r = r0 + j, kk = k0 + tx;
    if (r < R && kk < Kp) {
      float h, l;
      tf32_split(t[j][tx], h, l);
      hi[r * Kp + kk] = h;
      lo[r * Kp + kk] = l;
    }
  }
}

// ---------------- product ------------------------------------------------------------------------------------
struct Tf32Params {
  int64_t m, n, k;
  float* C;
  int64_t ldc;
  float alpha, beta;
  int beta_zero;
  int64_t tiles_m, tiles_n;
  int kchunk;  // k-blocks accumulated in TMEM before the partial sum is folded into C (see KCHUNK_DEFAULT)
  int uplo;    // triangular mask of ?syrk_ (common.cuh)
};

// The tensor core adds into the fp32 TMEM accumulator without round-to-nearest; over a long k the truncation
// bias grows linearly with the chain length (measured, profiles/accuracy_tf32x3_r01.md: relative Frobenius error
// 15 eps per 256 k of uninterrupted chain, 963 eps and gauge ratio 54 at k = 16384).  The chain is therefore cut every
// KCHUNK_DEFAULT*32 k values: drain warps pull each partial sum out of TMEM and add it with IEEE fp32 adds into
// register accumulators while the next chunk accumulates in the other TMEM buffer -- the analogue of the reference's
// kc blocking, where every kc block ends in one rounding into C (GeneralBlockPanelKernel.h:1025-1066), without
// touching C more than once per tile.
constexpr int KCHUNK_DEFAULT = 4;

constexpr int GROUP = 16;  // tile rasterisation: a wave of 148 tiles covers ~16 x 9 tiles (2048 x 2304 of C): balanced A/B panel reuse in L2
__device__ __forceinline__ void tile_of(int64_t pid, int64_t tiles_m, int64_t tiles_n, int64_t& tm, int64_t& tn) {
  const int64_t per_group = GROUP * tiles_n;
  const int64_t g = pid / per_group;
  const int64_t first = g * GROUP;
  const int64_t gsz = (tiles_m - first < GROUP) ? (tiles_m - first) : GROUP;
  const int64_t in = pid - g * per_group;
  tm = first + in % gsz;
  tn = in / gsz;
}

__global__ void __launch_bounds__(THREADS_R, 1)
tf32x3_gemm_kernel(const __grid_constant__ CUtensorMap mapAh, const __grid_constant__ CUtensorMap mapAl,
                   const __grid_constant__ CUtensorMap mapBh, const __grid_constant__ CUtensorMap mapBl,
                   const Tf32Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024-byte alignment
  const uint32_t bars = base + NSTAGE * STAGE_BYTES;
  // barrier map (8 bytes each): full[NSTAGE] | empty[NSTAGE] | tfull[2] | tempty[2] | tmem ptr (4 bytes)
  auto full = [&](int s) { return bars + 8u * s; };
  auto empty = [&](int s) { return bars + 8u * (NSTAGE + s); };
  auto tfull = [&](int a) { return bars + 8u * (2 * NSTAGE + a); };
  auto tempty = [&](int a) { return bars + 8u * (2 * NSTAGE + 2 + a); };
  const uint32_t tmem_slot = bars + 8u * (2 * NSTAGE + 4);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapAh)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapAl)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapBh)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapBl)) : "memory");
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull(a), 1); mbar_init(tempty(a), 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int64_t ntiles = p.tiles_m * p.tiles_n;
  const int nkb = (int)((p.k + TK - 1) / TK);

  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer =====
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        int64_t tm, tn;
        tile_of(tile, p.tiles_m, p.tiles_n, tm, tn);
        if (tile_outside(p.uplo, tm * TM, tm * TM + TM, tn * TN, tn * TN + TN)) continue;   // rank-k update: other triangle
        const int row_a = (int)(tm * TM), row_b = (int)(tn * TN);
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(empty(stage), phase ^ 1u);
          mbar_expect_tx(full(stage), STAGE_BYTES);
          const uint32_t s0 = base + stage * STAGE_BYTES;
          tma_load_2d(s0, &mapAh, kb * TK, row_a, full(stage));
          tma_load_2d(s0 + A_PLANE, &mapAl, kb * TK, row_a, full(stage));
          tma_load_2d(s0 + 2 * A_PLANE, &mapBh, kb * TK, row_b, full(stage));
          tma_load_2d(s0 + 2 * A_PLANE + B_PLANE, &mapBl, kb * TK, row_b, full(stage));
          if (++stage == NSTAGE) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ===== MMA issuer (single thread) =====
      constexpr uint32_t idesc = umma_idesc_tf32(TM, TN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        int64_t tm, tn;
        tile_of(tile, p.tiles_m, p.tiles_n, tm, tn);
        if (tile_outside(p.uplo, tm * TM, tm * TM + TM, tn * TN, tn * TN + TN)) continue;   // rank-k update: other triangle
        for (int kb0 = 0; kb0 < nkb; kb0 += p.kchunk) {   // one accumulation chunk per TMEM buffer
          mbar_wait(tempty(acc), acc_phase ^ 1u);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * TN);
          const int kb1 = kb0 + p.kchunk < nkb ? kb0 + p.kchunk : nkb;
          for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(full(stage), phase);
            tc_fence_after();
            const uint32_t s0 = base + stage * STAGE_BYTES;
            const uint64_t a_hi = umma_desc_sw128(s0), a_lo = umma_desc_sw128(s0 + A_PLANE);
            const uint64_t b_hi = umma_desc_sw128(s0 + 2 * A_PLANE), b_lo = umma_desc_sw128(s0 + 2 * A_PLANE + B_PLANE);
#pragma unroll
            for (int k8 = 0; k8 < TK / UK; ++k8) {
              const uint64_t off = (uint64_t)((k8 * UK * 4) >> 4);  // advance inside the 128-byte swizzle atom
              tc_mma_tf32(d_tmem, a_lo + off, b_hi + off, idesc, ((kb - kb0) | k8) != 0);
              tc_mma_tf32(d_tmem, a_hi + off, b_lo + off, idesc, 1u);
              tc_mma_tf32(d_tmem, a_hi + off, b_hi + off, idesc, 1u);
            }
            tc_commit(empty(stage));  // frees the smem stage once these MMAs have read it
            if (++stage == NSTAGE) { stage = 0; phase ^= 1u; }
          }
          tc_commit(tfull(acc));  // partial sum complete -> epilogue
          if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
      }
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    // ===== drain warps 4..11: TMEM partial sums -> fp32 register accumulators -> C (alpha/beta fused) =====
    // The tensor core adds into TMEM without round-to-nearest, so the chain is cut every kchunk k-blocks: each
    // partial sum is pulled out with tcgen05.ld and added (IEEE fp32, round to nearest) into 128 register
    // accumulators per thread while the next chunk accumulates in the other TMEM buffer.  C is touched once per tile.
    const int q = warp & 3;              // TMEM lane quadrant this warp may access
    const int half = (warp - 4) >> 2;    // which 128 of the 256 accumulator columns
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      int64_t tm, tn;
      tile_of(tile, p.tiles_m, p.tiles_n, tm, tn);
        if (tile_outside(p.uplo, tm * TM, tm * TM + TM, tn * TN, tn * TN + TN)) continue;   // rank-k update: other triangle
      float accr[128];
#pragma unroll
      for (int j = 0; j < 128; ++j) accr[j] = 0.f;
      for (int kb0 = 0; kb0 < nkb; kb0 += p.kchunk) {
        mbar_wait(tfull(acc), acc_phase);
        tc_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < 128; c0 += 16) {
          uint32_t r[16];
          tmem_ld16(tmem_base + (uint32_t)(acc * TN + half * 128 + c0) + ((uint32_t)(q * 32) << 16), r);
#pragma unroll
          for (int j = 0; j < 16; ++j) accr[c0 + j] += __uint_as_float(r[j]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
      const int64_t row = tm * TM + q * 32 + lane;
      if (row < p.m) {
        const int64_t colbase = tn * TN + half * 128;
        float* pc = p.C + row + colbase * p.ldc;
        const int ncol = p.n - colbase > 128 ? 128 : (int)(p.n - colbase);   // valid columns of this half (may be <= 0)
#pragma unroll
        for (int c0 = 0; c0 < 128; c0 += 8) {
          // loads of 8 columns first, then their stores: a load placed after a store through the same pointer
          // cannot be hoisted by the compiler, which would serialise every memory round trip
          float old[8];
#pragma unroll
          for (int j = 0; j < 8; ++j)
            old[j] = (!p.beta_zero && c0 + j < ncol && in_triangle(p.uplo, row, colbase + c0 + j)) ? pc[(int64_t)j * p.ldc] : 0.f;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (c0 + j < ncol && in_triangle(p.uplo, row, colbase + c0 + j))
              pc[(int64_t)j * p.ldc] = fmaf(p.beta, old[j], p.alpha * accr[c0 + j]);
          pc += 8 * p.ldc;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// ---------------- 256x256 tile variant (large k): 1.5x less L2->SM traffic per flop ------------------------------
// The 128x256 kernel above needs 64 B/clk/SM of operands at full tensor rate, more than L2 delivers (~42 B/clk/SM):
// ncu shows it operand-starved at ~65-78 % tensor-pipe activity.  Here one CTA owns a 256x256 tile of C as two
// 128x256 accumulators that fill all 512 TMEM columns (no TMEM double buffering: the epilogue of a tile is < 3 % of
// its main loop once k >= 2048), every B k-block is used by both halves, and the stage shrinks to
// TK2 = 16 floats (64-byte rows, SWIZZLE_64B) so that three 64 KB stages fit: 42 B/clk/SM.
constexpr int TM2 = 256, TK2 = 16, NSTAGE2 = 3;
constexpr int PLANE2 = 256 * TK2 * 4;                 // 16384: one 256-row hi or lo plane of A or B
constexpr int STAGE2_BYTES = 4 * PLANE2;              // 65536
constexpr int SMEM2_BYTES = NSTAGE2 * STAGE2_BYTES + 1024 + 256;

// K-major operand, 64-byte rows, SWIZZLE_64B: 8-row groups are 512 bytes apart
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;  // SWIZZLE_64B
  return d;
}

__global__ void __launch_bounds__(THREADS, 1)
tf32x3_gemm256_kernel(const __grid_constant__ CUtensorMap mapAh, const __grid_constant__ CUtensorMap mapAl,
                      const __grid_constant__ CUtensorMap mapBh, const __grid_constant__ CUtensorMap mapBl,
                      const Tf32Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t bars = base + NSTAGE2 * STAGE2_BYTES;
  auto full = [&](int s) { return bars + 8u * s; };
  auto empty = [&](int s) { return bars + 8u * (NSTAGE2 + s); };
  const uint32_t tfull = bars + 8u * (2 * NSTAGE2), tempty = bars + 8u * (2 * NSTAGE2 + 1);
  const uint32_t tmem_slot = bars + 8u * (2 * NSTAGE2 + 2);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapAh)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapAl)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapBh)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapBl)) : "memory");
    for (int s = 0; s < NSTAGE2; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
    mbar_init(tfull, 1);
    mbar_init(tempty, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const int64_t ntiles = p.tiles_m * p.tiles_n;
  const int nkb = (int)((p.k + TK2 - 1) / TK2);

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer =====
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        int64_t tm, tn;
        tile_of(tile, p.tiles_m, p.tiles_n, tm, tn);
        const int row_a = (int)(tm * TM2), row_b = (int)(tn * TN);
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(empty(stage), phase ^ 1u);
          mbar_expect_tx(full(stage), STAGE2_BYTES);
          const uint32_t s0 = base + stage * STAGE2_BYTES;
          tma_load_2d(s0, &mapAh, kb * TK2, row_a, full(stage));
          tma_load_2d(s0 + PLANE2, &mapAl, kb * TK2, row_a, full(stage));
          tma_load_2d(s0 + 2 * PLANE2, &mapBh, kb * TK2, row_b, full(stage));
          tma_load_2d(s0 + 3 * PLANE2, &mapBl, kb * TK2, row_b, full(stage));
          if (++stage == NSTAGE2) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ===== MMA issuer =====
      constexpr uint32_t idesc = umma_idesc_tf32(TM, TN);
      int stage = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        mbar_wait(tempty, acc_phase ^ 1u);  // the epilogue has drained the previous tile
        tc_fence_after();
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(full(stage), phase);
          tc_fence_after();
          const uint32_t s0 = base + stage * STAGE2_BYTES;
          const uint64_t b_hi = umma_desc_sw64(s0 + 2 * PLANE2), b_lo = umma_desc_sw64(s0 + 3 * PLANE2);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint64_t a_hi = umma_desc_sw64(s0 + h * (PLANE2 / 2)), a_lo = umma_desc_sw64(s0 + PLANE2 + h * (PLANE2 / 2));
            const uint32_t d_tmem = tmem_base + (uint32_t)(h * TN);
#pragma unroll
            for (int k8 = 0; k8 < TK2 / UK; ++k8) {
              const uint64_t off = (uint64_t)((k8 * UK * 4) >> 4);
              tc_mma_tf32(d_tmem, a_lo + off, b_hi + off, idesc, (kb | k8) != 0);
              tc_mma_tf32(d_tmem, a_hi + off, b_lo + off, idesc, 1u);
              tc_mma_tf32(d_tmem, a_hi + off, b_hi + off, idesc, 1u);
            }
          }
          tc_commit(empty(stage));
          if (++stage == NSTAGE2) { stage = 0; phase ^= 1u; }
        }
        tc_commit(tfull);
        acc_phase ^= 1u;
      }
    }
  } else {
    const int q = warp & 3;
    uint32_t acc_phase = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      int64_t tm, tn;
      tile_of(tile, p.tiles_m, p.tiles_n, tm, tn);
      mbar_wait(tfull, acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        const int64_t row = tm * TM2 + h * TM + q * 32 + lane;
        float* crow = p.C + row;
#pragma unroll 1
        for (int c0 = 0; c0 < TN; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(tmem_base + (uint32_t)(h * TN + c0) + ((uint32_t)(q * 32) << 16), r);
          const int64_t col0 = tn * TN + c0;
          if (row < p.m) {
            float old[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) old[j] = (!p.beta_zero && col0 + j < p.n) ? crow[(col0 + j) * p.ldc] : 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.n) crow[(col0 + j) * p.ldc] = fmaf(p.beta, old[j], p.alpha * __uint_as_float(r[j]));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty);
      acc_phase ^= 1u;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// ---------------- CTA-pair variant (cta_group::2): 256x256 tile per pair, half the B reads per SM --------------------
// The single-CTA kernel is limited by the shared-memory port: every SS-mode tcgen05.mma re-reads A (4 KB) and B (8 KB)
// from shared memory (96 B/clk) on top of the TMA fill (64 B/clk) against 128 B/clk.  A CTA pair (two SMs of one TPC)
// multiplies a 256x256 tile with M=256 MMAs issued by the leader CTA: each CTA stages its own 128 rows of A and only ITS
// half (128 rows) of B^T, the tensor cores exchange the B halves, so per SM the MMA reads 8 KB per instruction
// (64 B/clk) and the TMA fill is 64 KB per 1536 clk (42 B/clk): 106 B/clk.  Three 64 KB stages fit.
constexpr int P_STAGE_BYTES = 4 * A_PLANE;   // A_hi, A_lo, B_hi(half), B_lo(half): 4 x 16 KB
constexpr int P_NSTAGE = 3;
constexpr int P_SMEM_BYTES = P_NSTAGE * P_STAGE_BYTES + 1024 + 256;
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;  // shared::cluster address of the same offset in the pair's leader (even) CTA

__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t leader_bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(leader_bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive (once the MMAs issued so far are done) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS_R, 1)
tf32x3_gemm_pair_kernel(const __grid_constant__ CUtensorMap mapAh, const __grid_constant__ CUtensorMap mapAl,
                        const __grid_constant__ CUtensorMap mapBh, const __grid_constant__ CUtensorMap mapBl,
                        const Tf32Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t bars = base + P_NSTAGE * P_STAGE_BYTES;
  auto full = [&](int s) { return bars + 8u * s; };
  auto empty = [&](int s) { return bars + 8u * (P_NSTAGE + s); };
  auto tfull = [&](int a) { return bars + 8u * (2 * P_NSTAGE + a); };
  auto tempty = [&](int a) { return bars + 8u * (2 * P_NSTAGE + 2 + a); };
  const uint32_t tmem_slot = bars + 8u * (2 * P_NSTAGE + 4);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapAh)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapAl)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapBh)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mapBl)) : "memory");
    for (int s = 0; s < P_NSTAGE; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull(a), 1); mbar_init(tempty(a), 16); }  // 8 drain warps x 2 CTAs
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int64_t ntiles = p.tiles_m * p.tiles_n;   // 256 x 256 tiles
  const int nkb = (int)((p.k + TK - 1) / TK);
  const int64_t pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 0) {
      if (lane == 0) {  // ===== TMA producer (both CTAs; completion is counted on the LEADER's full barrier) =====
        int stage = 0;
        uint32_t phase = 0;
        for (int64_t tile = pair; tile < ntiles; tile += npairs) {
          int64_t tm, tn;
          tile_of(tile, p.tiles_m, p.tiles_n, tm, tn);
        if (tile_outside(p.uplo, tm * 256, tm * 256 + 256, tn * 256, tn * 256 + 256)) continue;   // rank-k update: other triangle
          const int row_a = (int)(tm * 256 + rank * 128), row_b = (int)(tn * 256 + rank * 128);
          for (int kb = 0; kb < nkb; ++kb) {
            mbar_wait(empty(stage), phase ^ 1u);
            if (leader) mbar_expect_tx(full(stage), 2 * P_STAGE_BYTES);
            const uint32_t s0 = base + stage * P_STAGE_BYTES;
            const uint32_t lbar = full(stage) & PEER_MASK;
            tma_load_2d_pair(s0, &mapAh, kb * TK, row_a, lbar);
            tma_load_2d_pair(s0 + A_PLANE, &mapAl, kb * TK, row_a, lbar);
            tma_load_2d_pair(s0 + 2 * A_PLANE, &mapBh, kb * TK, row_b, lbar);
            tma_load_2d_pair(s0 + 3 * A_PLANE, &mapBl, kb * TK, row_b, lbar);
            if (++stage == P_NSTAGE) { stage = 0; phase ^= 1u; }
          }
        }
      }
    } else if (warp == 1 && leader) {
      if (lane == 0) {  // ===== MMA issuer: leader CTA only, M = 256 across the pair =====
        constexpr uint32_t idesc = umma_idesc_tf32(256, 256);
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int64_t tile = pair; tile < ntiles; tile += npairs) {
          int64_t tm, tn;
          tile_of(tile, p.tiles_m, p.tiles_n, tm, tn);
          if (tile_outside(p.uplo, tm * 256, tm * 256 + 256, tn * 256, tn * 256 + 256)) continue;   // rank-k update: other triangle
          for (int kb0 = 0; kb0 < nkb; kb0 += p.kchunk) {
            mbar_wait(tempty(acc), acc_phase ^ 1u);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * TN);
            const int kb1 = kb0 + p.kchunk < nkb ? kb0 + p.kchunk : nkb;
            for (int kb = kb0; kb < kb1; ++kb) {
              mbar_wait(full(stage), phase);
              tc_fence_after();
              const uint32_t s0 = base + stage * P_STAGE_BYTES;
              const uint64_t a_hi = umma_desc_sw128(s0), a_lo = umma_desc_sw128(s0 + A_PLANE);
              const uint64_t b_hi = umma_desc_sw128(s0 + 2 * A_PLANE), b_lo = umma_desc_sw128(s0 + 3 * A_PLANE);
#pragma unroll
              for (int k8 = 0; k8 < TK / UK; ++k8) {
                const uint64_t off = (uint64_t)((k8 * UK * 4) >> 4);
                tc_mma_tf32_pair(d_tmem, a_lo + off, b_hi + off, idesc, ((kb - kb0) | k8) != 0);
                tc_mma_tf32_pair(d_tmem, a_hi + off, b_lo + off, idesc, 1u);
                tc_mma_tf32_pair(d_tmem, a_hi + off, b_hi + off, idesc, 1u);
              }
              tc_commit_pair(empty(stage));   // frees this stage in BOTH CTAs
              if (++stage == P_NSTAGE) { stage = 0; phase ^= 1u; }
            }
            tc_commit_pair(tfull(acc));       // partial sum complete in both CTAs' TMEM
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
          }
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    // ===== drain warps (both CTAs): own 128 rows of the pair tile; tempty is counted on the leader's barrier =====
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int64_t tile = pair; tile < ntiles; tile += npairs) {
      int64_t tm, tn;
      tile_of(tile, p.tiles_m, p.tiles_n, tm, tn);
        if (tile_outside(p.uplo, tm * 256, tm * 256 + 256, tn * 256, tn * 256 + 256)) continue;   // rank-k update: other triangle
      float accr[128];
#pragma unroll
      for (int j = 0; j < 128; ++j) accr[j] = 0.f;
      for (int kb0 = 0; kb0 < nkb; kb0 += p.kchunk) {
        mbar_wait(tfull(acc), acc_phase);
        tc_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < 128; c0 += 16) {
          uint32_t r[16];
          tmem_ld16(tmem_base + (uint32_t)(acc * TN + half * 128 + c0) + ((uint32_t)(q * 32) << 16), r);
#pragma unroll
          for (int j = 0; j < 16; ++j) accr[c0 + j] += __uint_as_float(r[j]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(tempty(acc) & PEER_MASK);
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
      const int64_t row = tm * 256 + rank * 128 + q * 32 + lane;
      if (row < p.m) {
        const int64_t colbase = tn * TN + half * 128;
        float* pc = p.C + row + colbase * p.ldc;
        const int ncol = p.n - colbase > 128 ? 128 : (int)(p.n - colbase);
#pragma unroll
        for (int c0 = 0; c0 < 128; c0 += 8) {
          float old[8];
#pragma unroll
          for (int j = 0; j < 8; ++j)
            old[j] = (!p.beta_zero && c0 + j < ncol && in_triangle(p.uplo, row, colbase + c0 + j)) ? pc[(int64_t)j * p.ldc] : 0.f;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (c0 + j < ncol && in_triangle(p.uplo, row, colbase + c0 + j))
              pc[(int64_t)j * p.ldc] = fmaf(p.beta, old[j], p.alpha * accr[c0 + j]);
          pc += 8 * p.ldc;
        }
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// ---------------- complex<float>: same machinery on de-interleaved planes ---------------------------------------
// C = A*B with A = Ar + i Ai, B = Br + i Bi:  Re = Ar.Br - Ai.Bi,  Im = Ar.Bi + Ai.Br  (the four real products of
// the reference's complex gebp, GeneralBlockPanelKernel.h:566-744), each as a 3xTF32 product; the minus sign is the
// negate-A bit of the instruction descriptor; conjugation is folded in at pack time (conj_if, BlasUtil.h:43-124).
// Tile: 128 complex rows x 64 complex columns; TMEM holds Re (64 cols) and Im (64 cols), double buffered.
constexpr int CTN = 64;
constexpr int CB_PLANE = CTN * TK * 4;                       // 8192
constexpr int CSTAGE_BYTES = 4 * A_PLANE + 4 * CB_PLANE;     // 98304
constexpr int CSMEM_BYTES = NSTAGE * CSTAGE_BYTES + 1024 + 256;
constexpr int CTMEM_COLS = 256;

__global__ void __launch_bounds__(256)
tf32_split_pack_cplx_kernel(const float2* __restrict__ src, int64_t sr, int64_t sk, int64_t R, int64_t K, int conj,
                            float* __restrict__ re_hi, float* __restrict__ re_lo, float* __restrict__ im_hi,
                            float* __restrict__ im_lo, int64_t Kp) {
  __shared__ float2 t[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t r0 = (int64_t)blockIdx.y * 32, k0 = (int64_t)blockIdx.x * 32;
  if (sk == 1) {
#pragma unroll
    for (int j = ty; j < 32; j += 8) {
      const int64_t r = r0 + j, kk = k0 + tx;
      t[j][tx] = (r < R && kk < K) ? src[r * sr + kk] : make_float2(0.f, 0.f);
    }
  } else {
#pragma unroll
    for (int j = ty; j < 32; j += 8) {
      const int64_t kk = k0 + j, r = r0 + tx;
      t[tx][j] = (r < R && kk < K) ? src[r + kk * sk] : make_float2(0.f, 0.f);
    }
  }
  __syncthreads();
#pragma unroll
  for (int j = ty; j < 32; j += 8) {
    const int64_t r = r0 + j, kk = k0 + tx;
    if (r < R && kk < Kp) {
      float2 x = t[j][tx];
      if (conj) x.y = -x.y;
      float h, l;
      tf32_split(x.x, h, l);
      re_hi[r * Kp + kk] = h;
      re_lo[r * Kp + kk] = l;
      tf32_split(x.y, h, l);
      im_hi[r * Kp + kk] = h;
      im_lo[r * Kp + kk] = l;
    }
  }
}

struct CMaps { CUtensorMap a[4], b[4]; };  // plane order: re_hi, re_lo, im_hi, im_lo
struct CTf32Params {
  int64_t m, n, k;
  float2* C;
  int64_t ldc;
  float2 alpha, beta;
  int beta_zero;
  int64_t tiles_m, tiles_n;
  int kchunk;
  int uplo, herm;   // triangular mask of ?syrk_/?herk_; herm: imaginary part of the diagonal stored as zero
};

__global__ void __launch_bounds__(THREADS, 1)
tf32x3_cgemm_kernel(const __grid_constant__ CMaps maps, const CTf32Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t bars = base + NSTAGE * CSTAGE_BYTES;
  auto full = [&](int s) { return bars + 8u * s; };
  auto empty = [&](int s) { return bars + 8u * (NSTAGE + s); };
  auto tfull = [&](int a) { return bars + 8u * (2 * NSTAGE + a); };
  auto tempty = [&](int a) { return bars + 8u * (2 * NSTAGE + 2 + a); };
  const uint32_t tmem_slot = bars + 8u * (2 * NSTAGE + 4);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 4; ++i) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.a[i])) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.b[i])) : "memory");
    }
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull(a), 1); mbar_init(tempty(a), 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(CTMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const int64_t ntiles = p.tiles_m * p.tiles_n;
  const int nkb = (int)((p.k + TK - 1) / TK);

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer: 8 planes per stage =====
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        int64_t tm, tn;
        tile_of(tile, p.tiles_m, p.tiles_n, tm, tn);
        if (tile_outside(p.uplo, tm * TM, tm * TM + TM, tn * CTN, tn * CTN + CTN)) continue;   // rank-k update: other triangle
        const int row_a = (int)(tm * TM), row_b = (int)(tn * CTN);
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(empty(stage), phase ^ 1u);
          mbar_expect_tx(full(stage), CSTAGE_BYTES);
          const uint32_t s0 = base + stage * CSTAGE_BYTES;
#pragma unroll
          for (int i = 0; i < 4; ++i) tma_load_2d(s0 + i * A_PLANE, &maps.a[i], kb * TK, row_a, full(stage));
#pragma unroll
          for (int i = 0; i < 4; ++i) tma_load_2d(s0 + 4 * A_PLANE + i * CB_PLANE, &maps.b[i], kb * TK, row_b, full(stage));
          if (++stage == NSTAGE) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ===== MMA issuer =====
      constexpr uint32_t idesc = umma_idesc_tf32(TM, CTN);
      constexpr uint32_t idesc_neg = idesc | (1u << 13);  // negate A
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        int64_t tm, tn;
        tile_of(tile, p.tiles_m, p.tiles_n, tm, tn);
        if (tile_outside(p.uplo, tm * TM, tm * TM + TM, tn * CTN, tn * CTN + CTN)) continue;   // rank-k update: other triangle
        for (int kb0 = 0; kb0 < nkb; kb0 += p.kchunk) {
          mbar_wait(tempty(acc), acc_phase ^ 1u);
          tc_fence_after();
          const uint32_t d_re = tmem_base + (uint32_t)(acc * 2 * CTN), d_im = d_re + CTN;
          const int kb1 = kb0 + p.kchunk < nkb ? kb0 + p.kchunk : nkb;
          for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(full(stage), phase);
            tc_fence_after();
            const uint32_t s0 = base + stage * CSTAGE_BYTES;
            uint64_t a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = umma_desc_sw128(s0 + i * A_PLANE); b[i] = umma_desc_sw128(s0 + 4 * A_PLANE + i * CB_PLANE); }
            // index: 0 = re_hi, 1 = re_lo, 2 = im_hi, 3 = im_lo
#pragma unroll
            for (int k8 = 0; k8 < TK / UK; ++k8) {
              const uint64_t off = (uint64_t)((k8 * UK * 4) >> 4);
              const uint32_t first = ((kb - kb0) | k8) != 0;
              // Re += Ar.Br
              tc_mma_tf32(d_re, a[1] + off, b[0] + off, idesc, first);
              tc_mma_tf32(d_re, a[0] + off, b[1] + off, idesc, 1u);
              tc_mma_tf32(d_re, a[0] + off, b[0] + off, idesc, 1u);
              // Re -= Ai.Bi
              tc_mma_tf32(d_re, a[3] + off, b[2] + off, idesc_neg, 1u);
              tc_mma_tf32(d_re, a[2] + off, b[3] + off, idesc_neg, 1u);
              tc_mma_tf32(d_re, a[2] + off, b[2] + off, idesc_neg, 1u);
              // Im += Ar.Bi
              tc_mma_tf32(d_im, a[1] + off, b[2] + off, idesc, first);
              tc_mma_tf32(d_im, a[0] + off, b[3] + off, idesc, 1u);
              tc_mma_tf32(d_im, a[0] + off, b[2] + off, idesc, 1u);
              // Im += Ai.Br
              tc_mma_tf32(d_im, a[3] + off, b[0] + off, idesc, 1u);
              tc_mma_tf32(d_im, a[2] + off, b[1] + off, idesc, 1u);
              tc_mma_tf32(d_im, a[2] + off, b[0] + off, idesc, 1u);
            }
            tc_commit(empty(stage));
            if (++stage == NSTAGE) { stage = 0; phase ^= 1u; }
          }
          tc_commit(tfull(acc));
          if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
      }
    }
  } else {
    // drain warps: partial Re / Im sums -> 64 + 64 fp32 register accumulators (IEEE adds), C touched once per tile
    const int q = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      int64_t tm, tn;
      tile_of(tile, p.tiles_m, p.tiles_n, tm, tn);
        if (tile_outside(p.uplo, tm * TM, tm * TM + TM, tn * CTN, tn * CTN + CTN)) continue;   // rank-k update: other triangle
      float are[CTN], aim[CTN];
#pragma unroll
      for (int j = 0; j < CTN; ++j) { are[j] = 0.f; aim[j] = 0.f; }
      for (int kb0 = 0; kb0 < nkb; kb0 += p.kchunk) {
        mbar_wait(tfull(acc), acc_phase);
        tc_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < CTN; c0 += 32) {
          uint32_t r[32];
          const uint32_t t0 = tmem_base + (uint32_t)(acc * 2 * CTN + c0) + ((uint32_t)(q * 32) << 16);
          tmem_ld32(t0, r);
#pragma unroll
          for (int j = 0; j < 32; ++j) are[c0 + j] += __uint_as_float(r[j]);
          tmem_ld32(t0 + CTN, r);
#pragma unroll
          for (int j = 0; j < 32; ++j) aim[c0 + j] += __uint_as_float(r[j]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
      const int64_t row = tm * TM + q * 32 + lane;
      if (row < p.m) {
        float2* crow = p.C + row;
        const int64_t colbase = tn * CTN;
#pragma unroll
        for (int c0 = 0; c0 < CTN; c0 += 8) {
          // loads first, stores afterwards (a load behind a store to the same array cannot be hoisted)
          float2 old[8];
#pragma unroll
          for (int j = 0; j < 8; ++j)
            old[j] = (!p.beta_zero && colbase + c0 + j < p.n && in_triangle(p.uplo, row, colbase + c0 + j))
                         ? crow[(colbase + c0 + j) * p.ldc] : make_float2(0.f, 0.f);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (colbase + c0 + j < p.n && in_triangle(p.uplo, row, colbase + c0 + j)) {
              const float xr = are[c0 + j], xi = aim[c0 + j];
              float2 v = make_float2(fmaf(p.alpha.x, xr, -p.alpha.y * xi), fmaf(p.alpha.x, xi, p.alpha.y * xr));
              v.x += fmaf(p.beta.x, old[j].x, -p.beta.y * old[j].y);
              v.y += fmaf(p.beta.x, old[j].y, p.beta.y * old[j].x);
              if (p.herm && row == colbase + c0 + j) v.y = 0.f;
              crow[(colbase + c0 + j) * p.ldc] = v;
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(CTMEM_COLS) : "memory");
  }
}

// ---------------- complex CTA-pair variant: 256 x 128 complex tile per pair ------------------------------------------
// Same pairing as tf32x3_gemm_pair_kernel.  Per CTA and stage: 4 A planes of 128 rows (64 KB) + 4 B planes of 64 rows
// (32 KB), two stages.  M = 256, N = 128 MMAs; TMEM per buffer: Re (128 columns) | Im (128 columns), double buffered.
constexpr int CP_BN = 128;                                   // complex columns per pair tile
constexpr int CP_BPLANE = (CP_BN / 2) * TK * 4;              // 8192: this CTA's half of one B plane
constexpr int CP_STAGE_BYTES = 4 * A_PLANE + 4 * CP_BPLANE;  // 98304
constexpr int CP_SMEM_BYTES = NSTAGE * CP_STAGE_BYTES + 1024 + 256;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS_R, 1)
tf32x3_cgemm_pair_kernel(const __grid_constant__ CMaps maps, const CTf32Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t bars = base + NSTAGE * CP_STAGE_BYTES;
  auto full = [&](int s) { return bars + 8u * s; };
  auto empty = [&](int s) { return bars + 8u * (NSTAGE + s); };
  auto tfull = [&](int a) { return bars + 8u * (2 * NSTAGE + a); };
  auto tempty = [&](int a) { return bars + 8u * (2 * NSTAGE + 2 + a); };
  const uint32_t tmem_slot = bars + 8u * (2 * NSTAGE + 4);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 4; ++i) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.a[i])) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&maps.b[i])) : "memory");
    }
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull(a), 1); mbar_init(tempty(a), 16); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int64_t ntiles = p.tiles_m * p.tiles_n;   // 256 x 128 complex tiles
  const int nkb = (int)((p.k + TK - 1) / TK);
  const int64_t pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 0) {
      if (lane == 0) {  // ===== TMA producer =====
        int stage = 0;
        uint32_t phase = 0;
        for (int64_t tile = pair; tile < ntiles; tile += npairs) {
          int64_t tm, tn;
          tile_of(tile, p.tiles_m, p.tiles_n, tm, tn);
        if (tile_outside(p.uplo, tm * 256, tm * 256 + 256, tn * CP_BN, tn * CP_BN + CP_BN)) continue;   // rank-k update: other triangle
          const int row_a = (int)(tm * 256 + rank * 128), row_b = (int)(tn * CP_BN + rank * (CP_BN / 2));
          for (int kb = 0; kb < nkb; ++kb) {
            mbar_wait(empty(stage), phase ^ 1u);
            if (leader) mbar_expect_tx(full(stage), 2 * CP_STAGE_BYTES);
            const uint32_t s0 = base + stage * CP_STAGE_BYTES;
            const uint32_t lbar = full(stage) & PEER_MASK;
#pragma unroll
            for (int i = 0; i < 4; ++i) tma_load_2d_pair(s0 + i * A_PLANE, &maps.a[i], kb * TK, row_a, lbar);
#pragma unroll
            for (int i = 0; i < 4; ++i) tma_load_2d_pair(s0 + 4 * A_PLANE + i * CP_BPLANE, &maps.b[i], kb * TK, row_b, lbar);
            if (++stage == NSTAGE) { stage = 0; phase ^= 1u; }
          }
        }
      }
    } else if (warp == 1 && leader) {
      if (lane == 0) {  // ===== MMA issuer (leader CTA) =====
        constexpr uint32_t idesc = umma_idesc_tf32(256, CP_BN);
        constexpr uint32_t idesc_neg = idesc | (1u << 13);
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int64_t tile = pair; tile < ntiles; tile += npairs) {
          int64_t tm, tn;
          tile_of(tile, p.tiles_m, p.tiles_n, tm, tn);
          if (tile_outside(p.uplo, tm * 256, tm * 256 + 256, tn * CP_BN, tn * CP_BN + CP_BN)) continue;   // rank-k update: other triangle
          for (int kb0 = 0; kb0 < nkb; kb0 += p.kchunk) {
            mbar_wait(tempty(acc), acc_phase ^ 1u);
            tc_fence_after();
            const uint32_t d_re = tmem_base + (uint32_t)(acc * 2 * CP_BN), d_im = d_re + CP_BN;
            const int kb1 = kb0 + p.kchunk < nkb ? kb0 + p.kchunk : nkb;
            for (int kb = kb0; kb < kb1; ++kb) {
              mbar_wait(full(stage), phase);
              tc_fence_after();
              const uint32_t s0 = base + stage * CP_STAGE_BYTES;
              uint64_t a[4], b[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) { a[i] = umma_desc_sw128(s0 + i * A_PLANE); b[i] = umma_desc_sw128(s0 + 4 * A_PLANE + i * CP_BPLANE); }
#pragma unroll
              for (int k8 = 0; k8 < TK / UK; ++k8) {
                const uint64_t off = (uint64_t)((k8 * UK * 4) >> 4);
                const uint32_t first = ((kb - kb0) | k8) != 0;
                tc_mma_tf32_pair(d_re, a[1] + off, b[0] + off, idesc, first);
                tc_mma_tf32_pair(d_re, a[0] + off, b[1] + off, idesc, 1u);
                tc_mma_tf32_pair(d_re, a[0] + off, b[0] + off, idesc, 1u);
                tc_mma_tf32_pair(d_re, a[3] + off, b[2] + off, idesc_neg, 1u);
                tc_mma_tf32_pair(d_re, a[2] + off, b[3] + off, idesc_neg, 1u);
                tc_mma_tf32_pair(d_re, a[2] + off, b[2] + off, idesc_neg, 1u);
                tc_mma_tf32_pair(d_im, a[1] + off, b[2] + off, idesc, first);
                tc_mma_tf32_pair(d_im, a[0] + off, b[3] + off, idesc, 1u);
                tc_mma_tf32_pair(d_im, a[0] + off, b[2] + off, idesc, 1u);
                tc_mma_tf32_pair(d_im, a[3] + off, b[0] + off, idesc, 1u);
                tc_mma_tf32_pair(d_im, a[2] + off, b[1] + off, idesc, 1u);
                tc_mma_tf32_pair(d_im, a[2] + off, b[0] + off, idesc, 1u);
              }
              tc_commit_pair(empty(stage));
              if (++stage == NSTAGE) { stage = 0; phase ^= 1u; }
            }
            tc_commit_pair(tfull(acc));
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
          }
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    // drain warps: quadrant q, column half h -> Re[h*64, +64) and Im[h*64, +64) of this CTA's 128 rows
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int64_t tile = pair; tile < ntiles; tile += npairs) {
      int64_t tm, tn;
      tile_of(tile, p.tiles_m, p.tiles_n, tm, tn);
        if (tile_outside(p.uplo, tm * 256, tm * 256 + 256, tn * CP_BN, tn * CP_BN + CP_BN)) continue;   // rank-k update: other triangle
      float are[64], aim[64];
#pragma unroll
      for (int j = 0; j < 64; ++j) { are[j] = 0.f; aim[j] = 0.f; }
      for (int kb0 = 0; kb0 < nkb; kb0 += p.kchunk) {
        mbar_wait(tfull(acc), acc_phase);
        tc_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 16) {
          uint32_t r[16];
          const uint32_t t0 = tmem_base + (uint32_t)(acc * 2 * CP_BN + half * 64 + c0) + ((uint32_t)(q * 32) << 16);
          tmem_ld16(t0, r);
#pragma unroll
          for (int j = 0; j < 16; ++j) are[c0 + j] += __uint_as_float(r[j]);
          tmem_ld16(t0 + CP_BN, r);
#pragma unroll
          for (int j = 0; j < 16; ++j) aim[c0 + j] += __uint_as_float(r[j]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(tempty(acc) & PEER_MASK);
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
      const int64_t row = tm * 256 + rank * 128 + q * 32 + lane;
      if (row < p.m) {
        const int64_t colbase = tn * CP_BN + half * 64;
        float2* pc = p.C + row + colbase * p.ldc;
        const int ncol = p.n - colbase > 64 ? 64 : (int)(p.n - colbase);
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 8) {
          float2 old[8];
#pragma unroll
          for (int j = 0; j < 8; ++j)
            old[j] = (!p.beta_zero && c0 + j < ncol && in_triangle(p.uplo, row, colbase + c0 + j)) ? pc[(int64_t)j * p.ldc] : make_float2(0.f, 0.f);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (c0 + j < ncol && in_triangle(p.uplo, row, colbase + c0 + j)) {
              const float xr = are[c0 + j], xi = aim[c0 + j];
              float2 v = make_float2(fmaf(p.alpha.x, xr, -p.alpha.y * xi), fmaf(p.alpha.x, xi, p.alpha.y * xr));
              v.x += fmaf(p.beta.x, old[j].x, -p.beta.y * old[j].y);
              v.y += fmaf(p.beta.x, old[j].y, p.beta.y * old[j].x);
              if (p.herm && row == colbase + c0 + j) v.y = 0.f;
              pc[(int64_t)j * p.ldc] = v;
            }
          }
          pc += 8 * p.ldc;
        }
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// ---------------- TF32 pipe peak: back-to-back M=128 N=256 K=8 MMAs on zeroed shared memory ---------------------
__global__ void __launch_bounds__(128, 1) tf32_peak_kernel(int iters, float* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t bar = base + A_PLANE + B_PLANE;
  const uint32_t tmem_slot = bar + 8;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - raw));
  for (int i = threadIdx.x; i < (A_PLANE + B_PLANE) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem_raw + (base - raw))[i] = 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy zero fill -> async-proxy MMA reads
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  if (warp == 0 && lane == 0) {
    constexpr uint32_t idesc = umma_idesc_tf32(TM, TN);
    const uint64_t a = umma_desc_sw128(base), b = umma_desc_sw128(base + A_PLANE);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k8 = 0; k8 < 4; ++k8) tc_mma_tf32(tmem_base, a + 2 * k8, b + 2 * k8, idesc, (i | k8) != 0);
    }
    tc_commit(bar);
    mbar_wait(bar, 0);
    tc_fence_after();
  }
  __syncthreads();
  if (warp == 2) {
    uint32_t r[32];
    tmem_ld32(tmem_base + ((uint32_t)(64) << 16), r);
    if (r[lane] == 0x12345678u) out[0] = 1.f;  // keep the accumulator observable
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256) : "memory");
  }
}

// ---------------- host side ----------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) return (EncodeTiledFn) nullptr;
    if (q != cudaDriverEntryPointSuccess) return (EncodeTiledFn) nullptr;
    return (EncodeTiledFn)p;
  }();
  return fn;
}
// K-major panel: rows x K floats, row pitch Kp floats; box = 32 floats (128 bytes) x box_rows, SWIZZLE_128B
int make_map(CUtensorMap* map, const float* base, int64_t rows, int64_t K, int64_t Kp, int box_rows, int box_k = TK) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return (int)cudaErrorNotSupported;
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)Kp * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_k, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, box_k == 16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
}

int64_t kpad(int64_t k) { return (k + 31) / 32 * 32; }
int kchunk_blocks() {
  static const int v = [] { const char* e = getenv("B200BLAS_TF32_KCHUNK"); const int x = e ? atoi(e) : 0; return x > 0 ? x : KCHUNK_DEFAULT; }();
  return v;
}
int sm_count() {
  static int sms = [] {
    int dev = 0, n = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n > 0 ? n : 148;
  }();
  return sms;
}

}  // namespace

bool tf32x3_supported(const GemmProblem& p) {
  if (p.type != TY_S && p.type != TY_C) return false;
  const uintptr_t am = p.type == TY_C ? 7 : 3;
  if (((uintptr_t)p.A & am) || ((uintptr_t)p.B & am) || ((uintptr_t)p.C & am)) return false;
  if (p.m <= 0 || p.n <= 0 || p.k <= 0) return false;
  if (p.m > 0x7fffff00LL || p.n > 0x7fffff00LL || p.k > 0x7fffff00LL) return false;
  return encode_fn() != nullptr;
}

size_t tf32x3_workspace_bytes(const GemmProblem& p) {
  const size_t planes = p.type == TY_C ? 4 : 2;
  return planes * (size_t)(p.m + p.n) * (size_t)kpad(p.k) * sizeof(float) + 1024;
}

static int launch_tf32x3_cplx(const GemmProblem& p, cudaStream_t s, float* ws) {
  const int64_t Kp = kpad(p.k);
  float* ap[4];
  float* bp[4];
  for (int i = 0; i < 4; ++i) ap[i] = ws + (int64_t)i * p.m * Kp;
  for (int i = 0; i < 4; ++i) bp[i] = ws + 4 * p.m * Kp + (int64_t)i * p.n * Kp;
  {
    dim3 grid((unsigned)(Kp / 32), (unsigned)((p.m + 31) / 32));
    if (p.opa == OP_N) tf32_split_pack_cplx_kernel<<<grid, 256, 0, s>>>((const float2*)p.A, 1, p.lda, p.m, p.k, 0, ap[0], ap[1], ap[2], ap[3], Kp);
    else tf32_split_pack_cplx_kernel<<<grid, 256, 0, s>>>((const float2*)p.A, p.lda, 1, p.m, p.k, p.opa == OP_C, ap[0], ap[1], ap[2], ap[3], Kp);
    count_launch();
  }
  {
    dim3 grid((unsigned)(Kp / 32), (unsigned)((p.n + 31) / 32));
    if (p.opb == OP_N) tf32_split_pack_cplx_kernel<<<grid, 256, 0, s>>>((const float2*)p.B, p.ldb, 1, p.n, p.k, 0, bp[0], bp[1], bp[2], bp[3], Kp);
    else tf32_split_pack_cplx_kernel<<<grid, 256, 0, s>>>((const float2*)p.B, 1, p.ldb, p.n, p.k, p.opb == OP_C, bp[0], bp[1], bp[2], bp[3], Kp);
    count_launch();
  }
  B200_CUDA_TRY(cudaGetLastError());
  static const int force_pair = [] { const char* e = getenv("B200BLAS_TF32_PAIR"); return e ? atoi(e) : -1; }();
  const int64_t ptiles = ((p.m + 255) / 256) * ((p.n + CP_BN - 1) / CP_BN);
  const bool use_pair = force_pair >= 0 ? force_pair != 0 : (p.m >= 512 && p.n >= 256 && ptiles >= 32);
  if (use_pair) {
    CMaps maps;
    for (int i = 0; i < 4; ++i) {
      if (make_map(&maps.a[i], ap[i], p.m, p.k, Kp, 128) || make_map(&maps.b[i], bp[i], p.n, p.k, Kp, CP_BN / 2)) return (int)cudaErrorInvalidValue;
    }
    CTf32Params prm;
    prm.m = p.m; prm.n = p.n; prm.k = p.k;
    prm.C = (float2*)p.C; prm.ldc = p.ldc;
    prm.alpha = make_float2((float)p.alpha[0], (float)p.alpha[1]);
    prm.beta = make_float2((float)p.beta[0], (float)p.beta[1]);
    prm.beta_zero = (p.beta[0] == 0.0 && p.beta[1] == 0.0);
    prm.tiles_m = (p.m + 255) / 256; prm.tiles_n = (p.n + CP_BN - 1) / CP_BN;
    prm.kchunk = kchunk_blocks();
    prm.uplo = p.uplo; prm.herm = p.herm;
    B200_SET_MAX_DYN_SMEM_ONCE(tf32x3_cgemm_pair_kernel, CP_SMEM_BYTES);
    int64_t pairs = sm_count() / 2;
    if (ptiles < pairs) pairs = ptiles;
    note_variant("tf32x3_tcgen05_c_pair_256x128x32");
    tf32x3_cgemm_pair_kernel<<<(unsigned)(2 * pairs), THREADS_R, CP_SMEM_BYTES, s>>>(maps, prm);
    count_launch();
    return (int)cudaGetLastError();
  }
  CMaps maps;
  for (int i = 0; i < 4; ++i) {
    if (make_map(&maps.a[i], ap[i], p.m, p.k, Kp, TM) || make_map(&maps.b[i], bp[i], p.n, p.k, Kp, CTN)) return (int)cudaErrorInvalidValue;
  }
  CTf32Params prm;
  prm.m = p.m; prm.n = p.n; prm.k = p.k;
  prm.C = (float2*)p.C; prm.ldc = p.ldc;
  prm.alpha = make_float2((float)p.alpha[0], (float)p.alpha[1]);
  prm.beta = make_float2((float)p.beta[0], (float)p.beta[1]);
  prm.beta_zero = (p.beta[0] == 0.0 && p.beta[1] == 0.0);
  prm.tiles_m = (p.m + TM - 1) / TM; prm.tiles_n = (p.n + CTN - 1) / CTN;
  prm.kchunk = kchunk_blocks();
    prm.uplo = p.uplo; prm.herm = p.herm;
  B200_SET_MAX_DYN_SMEM_ONCE(tf32x3_cgemm_kernel, CSMEM_BYTES);
  const int64_t ntiles = prm.tiles_m * prm.tiles_n;
  const unsigned grid = (unsigned)(ntiles < sm_count() ? ntiles : sm_count());
  note_variant("tf32x3_tcgen05_c_128x64x32");
  tf32x3_cgemm_kernel<<<grid, THREADS, CSMEM_BYTES, s>>>(maps, prm);
  count_launch();
  return (int)cudaGetLastError();
}

int launch_tf32x3(const GemmProblem& p, cudaStream_t s, void* workspace, size_t workspace_bytes) {
  if (!workspace || workspace_bytes < tf32x3_workspace_bytes(p)) return (int)cudaErrorInvalidValue;
  const int64_t Kp = kpad(p.k);
  float* ws = (float*)(((uintptr_t)workspace + 1023) & ~(uintptr_t)1023);
  if (p.type == TY_C) return launch_tf32x3_cplx(p, s, ws);
  float* Ah = ws;
  float* Al = Ah + p.m * Kp;
  float* Bh = Al + p.m * Kp;
  float* Bl = Bh + p.n * Kp;
  // pack A: logical (i, kk) at A[i + kk*lda] ('N') or A[kk + i*lda] ('T'/'C'; conj is the identity for float)
  {
    dim3 grid((unsigned)(Kp / 32), (unsigned)((p.m + 31) / 32));
    if (p.opa == OP_N) tf32_split_pack_kernel<<<grid, 256, 0, s>>>((const float*)p.A, 1, p.lda, p.m, p.k, Ah, Al, Kp);
    else tf32_split_pack_kernel<<<grid, 256, 0, s>>>((const float*)p.A, p.lda, 1, p.m, p.k, Ah, Al, Kp);
    count_launch();
  }
  // pack B^T: logical (j, kk) at B[kk + j*ldb] ('N') or B[j + kk*ldb] ('T'/'C')
  {
    dim3 grid((unsigned)(Kp / 32), (unsigned)((p.n + 31) / 32));
    if (p.opb == OP_N) tf32_split_pack_kernel<<<grid, 256, 0, s>>>((const float*)p.B, p.ldb, 1, p.n, p.k, Bh, Bl, Kp);
    else tf32_split_pack_kernel<<<grid, 256, 0, s>>>((const float*)p.B, 1, p.ldb, p.n, p.k, Bh, Bl, Kp);
    count_launch();
  }
  B200_CUDA_TRY(cudaGetLastError());
  static const int force_tile = [] { const char* e = getenv("B200BLAS_TF32_TILE"); return e ? atoi(e) : 0; }();
  const bool big = force_tile == 256 && p.uplo == 0;  // measured slower than the 128x256 kernel (profiles/variant_sweep_r01.md): opt-in only
  if (big) {
    CUtensorMap mAh, mAl, mBh, mBl;
    if (make_map(&mAh, Ah, p.m, p.k, Kp, 256, TK2) || make_map(&mAl, Al, p.m, p.k, Kp, 256, TK2) ||
        make_map(&mBh, Bh, p.n, p.k, Kp, 256, TK2) || make_map(&mBl, Bl, p.n, p.k, Kp, 256, TK2))
      return (int)cudaErrorInvalidValue;
    Tf32Params prm;
    prm.m = p.m; prm.n = p.n; prm.k = p.k;
    prm.C = (float*)p.C; prm.ldc = p.ldc;
    prm.alpha = (float)p.alpha[0]; prm.beta = (float)p.beta[0];
    prm.beta_zero = (p.beta[0] == 0.0);
    prm.tiles_m = (p.m + TM2 - 1) / TM2; prm.tiles_n = (p.n + TN - 1) / TN;
    prm.kchunk = 1 << 30;
    prm.uplo = 0;
    B200_SET_MAX_DYN_SMEM_ONCE(tf32x3_gemm256_kernel, SMEM2_BYTES);
    const int64_t nt = prm.tiles_m * prm.tiles_n;
    const unsigned g2 = (unsigned)(nt < sm_count() ? nt : sm_count());
    note_variant("tf32x3_tcgen05_256x256x16");
    tf32x3_gemm256_kernel<<<g2, THREADS, SMEM2_BYTES, s>>>(mAh, mAl, mBh, mBl, prm);
    count_launch();
    return (int)cudaGetLastError();
  }
  // CTA pairs win once there are enough 256x256 tiles to occupy the 74 pairs (profiles/variant_sweep_r01.md:
  // 8192^3 276 vs 252 TF, 2048^3 162 vs 147 TF); B200BLAS_TF32_PAIR=0/1 forces the choice
  static const int force_pair = [] { const char* e = getenv("B200BLAS_TF32_PAIR"); return e ? atoi(e) : -1; }();
  const int64_t tiles256 = ((p.m + 255) / 256) * ((p.n + 255) / 256);
  const bool use_pair = force_pair >= 0 ? force_pair != 0 : (p.m >= 512 && p.n >= 512 && tiles256 >= 32);
  if (use_pair) {
    // CTA pairs: every CTA loads 128-row boxes of A and of B^T
    CUtensorMap mAh, mAl, mBh, mBl;
    if (make_map(&mAh, Ah, p.m, p.k, Kp, 128) || make_map(&mAl, Al, p.m, p.k, Kp, 128) ||
        make_map(&mBh, Bh, p.n, p.k, Kp, 128) || make_map(&mBl, Bl, p.n, p.k, Kp, 128))
      return (int)cudaErrorInvalidValue;
    Tf32Params prm;
    prm.m = p.m; prm.n = p.n; prm.k = p.k;
    prm.C = (float*)p.C; prm.ldc = p.ldc;
    prm.alpha = (float)p.alpha[0]; prm.beta = (float)p.beta[0];
    prm.beta_zero = (p.beta[0] == 0.0);
    prm.tiles_m = (p.m + 255) / 256; prm.tiles_n = (p.n + 255) / 256;
    prm.kchunk = kchunk_blocks();
    prm.uplo = p.uplo;
    B200_SET_MAX_DYN_SMEM_ONCE(tf32x3_gemm_pair_kernel, P_SMEM_BYTES);
    const int64_t nt = prm.tiles_m * prm.tiles_n;
    int64_t pairs = sm_count() / 2;
    if (nt < pairs) pairs = nt;
    note_variant("tf32x3_tcgen05_pair_256x256x32");
    tf32x3_gemm_pair_kernel<<<(unsigned)(2 * pairs), THREADS_R, P_SMEM_BYTES, s>>>(mAh, mAl, mBh, mBl, prm);
    count_launch();
    return (int)cudaGetLastError();
  }
  CUtensorMap mAh, mAl, mBh, mBl;
  if (make_map(&mAh, Ah, p.m, p.k, Kp, TM) || make_map(&mAl, Al, p.m, p.k, Kp, TM) ||
      make_map(&mBh, Bh, p.n, p.k, Kp, TN) || make_map(&mBl, Bl, p.n, p.k, Kp, TN))
    return (int)cudaErrorInvalidValue;
  Tf32Params prm;
  prm.m = p.m; prm.n = p.n; prm.k = p.k;
  prm.C = (float*)p.C; prm.ldc = p.ldc;
  prm.alpha = (float)p.alpha[0]; prm.beta = (float)p.beta[0];
  prm.beta_zero = (p.beta[0] == 0.0);
  prm.tiles_m = (p.m + TM - 1) / TM; prm.tiles_n = (p.n + TN - 1) / TN;
  prm.kchunk = kchunk_blocks();
    prm.uplo = p.uplo;
  B200_SET_MAX_DYN_SMEM_ONCE(tf32x3_gemm_kernel, SMEM_BYTES);
  const int64_t ntiles = prm.tiles_m * prm.tiles_n;
  const unsigned grid = (unsigned)(ntiles < sm_count() ? ntiles : sm_count());  // persistent: one CTA per SM
  note_variant("tf32x3_tcgen05_128x256x32");
  tf32x3_gemm_kernel<<<grid, THREADS_R, SMEM_BYTES, s>>>(mAh, mAl, mBh, mBl, prm);
  count_launch();
  return (int)cudaGetLastError();
}

double tf32_pipe_peak(int millis) {
  const int smem = A_PLANE + B_PLANE + 1024 + 64;
  if (cudaFuncSetAttribute(tf32_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -1.0;
  float* out = nullptr;
  if (cudaMalloc(&out, 64) != cudaSuccess) return -1.0;
  const int iters = 4096, blocks = sm_count();
  const double flops = (double)blocks * iters * 4.0 * 2.0 * TM * TN * UK;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  tf32_peak_kernel<<<blocks, 128, smem>>>(iters, out);
  count_launch();
  if (cudaDeviceSynchronize() != cudaSuccess) { cudaGetLastError(); cudaFree(out); return -1.0; }
  double best = 0.0, total = 0.0;
  while (total < millis) {
    cudaEventRecord(e0);
    tf32_peak_kernel<<<blocks, 128, smem>>>(iters, out);
    count_launch();
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) { cudaGetLastError(); best = -1.0; break; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    total += ms;
    const double tf = flops / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(out);
  return best;
}

}  // namespace b200
