// eigen_b200/csrc/tri.cu -- triangular solve / multiply and symmetric products built on the GEMM kernels (sm_100a).
//
// SURVEY.md section 8 rows f2 (trsm) and f4 (trmm, symm/hemm).  Reference being replaced:
//   triangular_solve_matrix            Eigen/src/Core/products/TriangularSolverMatrix.h:41-335   (?trsm_, blas/level3_impl.h:78-178)
//   product_triangular_matrix_matrix   Eigen/src/Core/products/TriangularMatrixMatrix.h:89-383   (?trmm_, blas/level3_impl.h:183-284)
//   product_selfadjoint_matrix         Eigen/src/Core/products/SelfadjointMatrixMatrix.h:19-455  (?symm_/?hemm_, level3_impl.h:287-355,505-562)
// The reference blocks the triangle into small panels solved by substitution and pushes everything else through
// gebp_kernel.  Here the same split is recursive: a triangle of order d is cut at d1 (a power-of-two multiple of the
// leaf order), the off-diagonal block becomes ONE large product on the tensor-pipe GEMM kernels (gemm_dmma.cu /
// gemm_tf32x3.cu) and only leaves of order <= 128 (64 for complex double) are handled by a substitution kernel, one
// right-hand-side vector per thread, walked in register-resident sub-vectors of 32 elements.  Symmetric / Hermitian operands are expanded from the referenced triangle into a
// dense device image (what blas/level3_impl.h:324-341 does on the host for the complex case) and multiplied by the
// GEMM kernels.
#include <cstdlib>

#include "../../include/b200blas.h"
#include "common.cuh"
#include "scalar.cuh"

namespace b200 {
namespace {

// ---- leaf: S x = b (SOLVE) or x := S b (multiply) for every right-hand-side vector ---------------------------------------------
// S is the canonical matrix of order <= LB = NSUB * NB that multiplies a right-hand-side VECTOR from the left:
//   side = left :  S = op(A) block,      vectors = columns of B
//   side = right:  S = op(A)^T block,    vectors = rows of B          (X op(A) = B  <=>  op(A)^T X^T = B^T)
// Entries outside the referenced triangle are zero, a unit diagonal is one, and rows/columns past nb are the identity,
// so the unrolled NB-step recurrences are valid for every nb <= LB.  In SOLVE mode the diagonal holds reciprocals
// (the reference multiplies by 1/diag as well, TriangularSolverMatrix.h:118-121).
// One thread owns one right-hand-side vector and walks it in NSUB sub-vectors of NB elements held in registers:
// the contributions of the other sub-vectors (re-read from B: own earlier stores in a solve, still-original values in a
// product) are NB x NB register-blocked rank updates against the shared-memory block, then the NB x NB diagonal block is
// solved / applied by a fully unrolled recurrence.  One launch therefore replaces 2*NSUB-1 launches of the recursion.
constexpr int LEAF_THREADS = 256;   // one right-hand-side vector per thread

template <typename T, int NB, int NSUB, bool LOWER, bool SOLVE>
__global__ void __launch_bounds__(LEAF_THREADS)
tri_leaf_kernel(int left, int op, int uplo, int unit, int nb, int64_t nrhs, const T* __restrict__ A, int64_t lda,
                T* B, int64_t ldb) {
  constexpr int LB = NB * NSUB;
  extern __shared__ __align__(16) unsigned char leaf_smem[];
  T* S = reinterpret_cast<T*>(leaf_smem);   // S[i * LB + j]; every access is a warp-wide broadcast
  {
    // all loads of a batch are issued before any is consumed (a load-per-iteration loop costs one DRAM latency each);
    // p runs along the stored column of A so that the global reads are coalesced
    const bool swap = left ? (op != OP_N) : (op == OP_N);
    constexpr int BATCH = (LB * LB / LEAF_THREADS >= 16) ? 16 : 8;
    for (int base = 0; base < LB * LB; base += LEAF_THREADS * BATCH) {
      T vals[BATCH];
#pragma unroll
      for (int q = 0; q < BATCH; ++q) {
        const int idx = base + q * LEAF_THREADS + threadIdx.x;
        const int r = idx % LB, c = idx / LB;   // element of A
        const bool referenced = idx < LB * LB && r < nb && c < nb && ((r == c) ? !unit : (uplo == UPLO_UPPER ? r < c : r > c));
        vals[q] = referenced ? A[r + c * lda] : ((r == c) ? sc_one<T>() : Sc<T>::zero());
      }
#pragma unroll
      for (int q = 0; q < BATCH; ++q) {
        const int idx = base + q * LEAF_THREADS + threadIdx.x;
        if (idx >= LB * LB) continue;
        const int r = idx % LB, c = idx / LB;
        const bool referenced = r < nb && c < nb && ((r == c) ? !unit : (uplo == UPLO_UPPER ? r < c : r > c));
        T v = vals[q];
        if (referenced) {
          if (op == OP_C) v = Sc<T>::conj(v);
          if (SOLVE && r == c) v = sc_recip<T>(v);
        }
        const int i = swap ? c : r, j = swap ? r : c;   // position in S
        S[i * LB + j] = v;
      }
    }
  }
  __syncthreads();
  const int64_t v = (int64_t)blockIdx.x * LEAF_THREADS + threadIdx.x;
  if (v >= nrhs) return;
  T* bp = left ? B + v * ldb : B + v;
  const int64_t bs = left ? 1 : ldb;
  const int nsub = (nb + NB - 1) / NB;
  // sub-vector order: a lower solve / upper product goes top-down, an upper solve / lower product bottom-up
  constexpr bool ASCENDING = (SOLVE == LOWER);
  for (int step = 0; step < nsub; ++step) {
    const int b = ASCENDING ? step : nsub - 1 - step;
    const T* Sb = S + (b * NB) * LB;   // rows of sub-block b
    T x[NB];
#pragma unroll
    for (int i = 0; i < NB; ++i) x[i] = (b * NB + i < nb) ? bp[(int64_t)(b * NB + i) * bs] : Sc<T>::zero();
    if constexpr (!SOLVE) {
      // diagonal block first (in place; the off-diagonal terms below only add)
      const T* D = Sb + b * NB;
      if constexpr (LOWER) {
#pragma unroll
        for (int i = NB - 1; i >= 0; --i) {
          T acc = Sc<T>::mul(D[i * LB + i], x[i]);
#pragma unroll
          for (int j = 0; j < i; ++j) Sc<T>::fma(acc, D[i * LB + j], x[j]);
          x[i] = acc;
        }
      } else {
#pragma unroll
        for (int i = 0; i < NB; ++i) {
          T acc = Sc<T>::mul(D[i * LB + i], x[i]);
#pragma unroll
          for (int j = i + 1; j < NB; ++j) Sc<T>::fma(acc, D[i * LB + j], x[j]);
          x[i] = acc;
        }
      }
    }
    // the other sub-vectors this one depends on: c < b for a lower S, c > b for an upper S.  Their elements come from
    // global memory in chunks of 8; the next chunk is requested before the current one is consumed, so one load
    // latency is exposed per sub-vector pair instead of one per chunk.
    {
      const int c_lo = LOWER ? 0 : b + 1, c_hi = LOWER ? b : nsub;
      constexpr int CH = 8;
      T cur[CH], nxt[CH];
      int c = c_lo, j0 = 0;
      bool have = c < c_hi;
      if (have) {
#pragma unroll
        for (int q = 0; q < CH; ++q) cur[q] = (c * NB + j0 + q < nb) ? bp[(int64_t)(c * NB + j0 + q) * bs] : Sc<T>::zero();
      }
      while (have) {
        int c2 = c, j2 = j0 + CH;
        if (j2 == NB) { j2 = 0; ++c2; }
        const bool have2 = c2 < c_hi;
        if (have2) {
#pragma unroll
          for (int q = 0; q < CH; ++q) nxt[q] = (c2 * NB + j2 + q < nb) ? bp[(int64_t)(c2 * NB + j2 + q) * bs] : Sc<T>::zero();
        }
        const T* Sc_ = Sb + c * NB + j0;
#pragma unroll
        for (int q = 0; q < CH; ++q) {
#pragma unroll
          for (int i = 0; i < NB; ++i) {
            if constexpr (SOLVE) sc_fnma<T>(x[i], Sc_[i * LB + q], cur[q]);
            else Sc<T>::fma(x[i], Sc_[i * LB + q], cur[q]);
          }
        }
#pragma unroll
        for (int q = 0; q < CH; ++q) cur[q] = nxt[q];
        c = c2; j0 = j2; have = have2;
      }
    }
    if constexpr (SOLVE) {
      const T* D = Sb + b * NB;
      if constexpr (LOWER) {
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          x[j] = Sc<T>::mul(x[j], D[j * LB + j]);
#pragma unroll
          for (int i = j + 1; i < NB; ++i) sc_fnma<T>(x[i], D[i * LB + j], x[j]);
        }
      } else {
#pragma unroll
        for (int j = NB - 1; j >= 0; --j) {
          x[j] = Sc<T>::mul(x[j], D[j * LB + j]);
#pragma unroll
          for (int i = 0; i < j; ++i) sc_fnma<T>(x[i], D[i * LB + j], x[j]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < NB; ++i)
      if (b * NB + i < nb) bp[(int64_t)(b * NB + i) * bs] = x[i];
  }
}

// ---- block solve: S x = b for a LOWER canonical S of order <= LB, right-hand sides staged in shared memory -------------------------
// The substitution leaf above gives every thread one whole right-hand side (128 dependent steps per thread, strided global
// accesses for side = left: 42 - 70 us per launch).  Here a CTA owns NV right-hand sides as a tile X[i][v] in shared memory and
// walks the triangle in blocks of 32 rows: (1) the rows of the block are updated with the already solved part by ALL threads
// (4 x 2 register tiles, S broadcast from shared memory), (2) the 32 x 32 diagonal block is solved by one thread per
// right-hand side with the row ROTATED through registers (x[0] is always the next unknown; a compact runtime loop instead of
// 500 unrolled FMAs).  Used for every solve whose canonical matrix is lower (left: lower/N, upper/T,C; right: upper/N,
// lower/T,C) -- in particular the panel solves of ?potrf_ and the row solves of ?getrf_.
template <typename T, int LB, int NV>
__global__ void __launch_bounds__(256)
tri_block_solve_kernel(int left, int op, int uplo, int unit, int nb, int64_t nrhs, const T* __restrict__ A, int64_t lda,
                       T* B, int64_t ldb) {
  constexpr int LDSS = LB + 1, LDX = NV + 1;
  extern __shared__ __align__(16) unsigned char bs_smem[];
  T* S = reinterpret_cast<T*>(bs_smem);   // S[i * LDSS + j], lower canonical, reciprocal diagonal
  T* X = S + LB * LDSS;                    // X[i * LDX + v]
  const int tid = threadIdx.x;
  const bool swap = left ? (op != OP_N) : (op == OP_N);
  const int64_t v0 = (int64_t)blockIdx.x * NV;
  const int nv = (int)min((int64_t)NV, nrhs - v0);
  {
    // all loads of a batch are issued before any is consumed (a load-per-iteration loop costs one DRAM latency per element)
    constexpr int BATCH = 16;
    for (int base = 0; base < LB * LB; base += 256 * BATCH) {
      T vals[BATCH];
#pragma unroll
      for (int q = 0; q < BATCH; ++q) {
        const int idx = base + q * 256 + tid;
        const int r = idx % LB, c = idx / LB;   // element of A (coalesced along r)
        const bool referenced = idx < LB * LB && r < nb && c < nb && ((r == c) ? !unit : (uplo == UPLO_UPPER ? r < c : r > c));
        vals[q] = (r == c) ? sc_one<T>() : Sc<T>::zero();
        if (referenced) vals[q] = A[r + (int64_t)c * lda];
      }
#pragma unroll
      for (int q = 0; q < BATCH; ++q) {
        const int idx = base + q * 256 + tid;
        if (idx >= LB * LB) continue;
        const int r = idx % LB, c = idx / LB;
        const bool referenced = r < nb && c < nb && ((r == c) ? !unit : (uplo == UPLO_UPPER ? r < c : r > c));
        T v = vals[q];
        if (referenced) {
          if (op == OP_C) v = Sc<T>::conj(v);
          if (r == c) v = sc_recip<T>(v);
        }
        const int i = swap ? c : r, j = swap ? r : c;
        S[i * LDSS + j] = v;
      }
    }
    for (int base = 0; base < LB * NV; base += 256 * BATCH) {
      T vals[BATCH];
#pragma unroll
      for (int q = 0; q < BATCH; ++q) {
        const int idx = base + q * 256 + tid;
        int i, v;
        if (left) { i = idx % LB; v = idx / LB; } else { v = idx % NV; i = idx / NV; }
        vals[q] = Sc<T>::zero();
        if (idx < LB * NV && i < nb && v < nv) vals[q] = left ? B[i + (v0 + v) * ldb] : B[(v0 + v) + (int64_t)i * ldb];
      }
#pragma unroll
      for (int q = 0; q < BATCH; ++q) {
        const int idx = base + q * 256 + tid;
        int i, v;
        if (left) { i = idx % LB; v = idx / LB; } else { v = idx % NV; i = idx / NV; }
        if (idx < LB * NV) X[i * LDX + v] = vals[q];
      }
    }
  }
  __syncthreads();
  const int nblk = (nb + 31) / 32;
  for (int b = 0; b < nblk; ++b) {
    const int r0 = 32 * b;
    if (b > 0) {
      // (1) X[r0 + ii][v] -= sum_{q < r0} S[r0 + ii][q] X[q][v]: 32 x NV outputs, 4 rows x (NV / 32) vectors per thread
      constexpr int VT = NV / 32;
      const int tr = (tid / 32) * 4, tv = tid % 32;   // rows r0 + tr .. + 3, vectors tv + 32 * e
      T acc[4][VT];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int e = 0; e < VT; ++e) acc[r][e] = X[(r0 + tr + r) * LDX + tv + 32 * e];
      const T* srow = S + (r0 + tr) * LDSS;
#pragma unroll 4
      for (int q = 0; q < r0; ++q) {
        T sv[4], xv[VT];
#pragma unroll
        for (int r = 0; r < 4; ++r) sv[r] = srow[r * LDSS + q];          // warp-wide broadcasts
#pragma unroll
        for (int e = 0; e < VT; ++e) xv[e] = X[q * LDX + tv + 32 * e];  // consecutive lanes, consecutive words
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int e = 0; e < VT; ++e) sc_fnma<T>(acc[r][e], sv[r], xv[e]);
      }
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int e = 0; e < VT; ++e) X[(r0 + tr + r) * LDX + tv + 32 * e] = acc[r][e];
      __syncthreads();
    }
    // (2) diagonal block: one thread per right-hand side, rotated row
    if (tid < NV) {
      T x[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) x[i] = X[(r0 + i) * LDX + tid];
#pragma unroll 1
      for (int q = 0; q < 32; ++q) {
        const T* scol = S + (r0 + q) * LDSS + r0 + q;   // S[r0 + q + i][r0 + q] = scol[i * LDSS]
        const T xq = Sc<T>::mul(x[0], scol[0]);        // reciprocal diagonal (1 for unit / identity rows)
        X[(r0 + q) * LDX + tid] = xq;
#pragma unroll
        for (int i = 1; i < 32; ++i) {
          T v = x[i];
          if (q + i < 32) sc_fnma<T>(v, scol[i * LDSS], xq);
          x[i - 1] = v;
        }
      }
    }
    __syncthreads();
  }
  if (left) {
    for (int idx = tid; idx < LB * NV; idx += 256) {
      const int i = idx % LB, v = idx / LB;
      if (i < nb && v < nv) B[i + (v0 + v) * ldb] = X[i * LDX + v];
    }
  } else {
    for (int idx = tid; idx < LB * NV; idx += 256) {
      const int v = idx % NV, i = idx / NV;
      if (i < nb && v < nv) B[(v0 + v) + (int64_t)i * ldb] = X[i * LDX + v];
    }
  }
}

// ---- inverse-based leaves (default for large UPPER canonical solves, B200BLAS_TRSM=inv forces them everywhere) --------------------
// Inverses of all diagonal leaf blocks of T = op(A) in ONE launch (one CTA per block), so that a leaf of the solve is a
// product X_b = inv(T_bb) * B_b on the tensor-pipe kernels instead of 8256 FMA + LDS per right-hand side on the SIMT
// pipe (profiles/launches_r01_dpotrf8192_v2.md).  The block is brought to lower-canonical form Lc (T itself, or T^T
// for an upper T), split into halves [L11 0; L21 L22]; threads 0..2H-1 invert L11 and L22 one column each (registers,
// broadcast reads of the block), then all threads form X21 = -X22 * (L21 * X11); V = X (or X^T for an upper T).
template <typename T, int IB>
__global__ void __launch_bounds__(256)
trtri_diag_kernel(int op, int uplo, int unit, int t_lower, int64_t n, const T* __restrict__ A, int64_t lda, T* __restrict__ V) {
  constexpr int H = IB / 2;
  extern __shared__ __align__(16) unsigned char inv_smem[];
  T* L11 = reinterpret_cast<T*>(inv_smem);   // [i * H + j]
  T* L22 = L11 + H * H;
  T* L21 = L22 + H * H;
  T* X11 = L21 + H * H;
  T* X22 = X11 + H * H;
  const int tid = threadIdx.x;
  const int64_t d0 = (int64_t)blockIdx.x * IB;
  const int nb = (int)min((int64_t)IB, n - d0);
  const T* Ab = A + d0 + d0 * lda;
  T* Vb = V + (int64_t)blockIdx.x * IB * IB;
  for (int idx = tid; idx < IB * IB; idx += 256) {
    const int i = idx % IB, j = idx / IB;   // element (i, j) of Lc
    T v = (i == j) ? sc_one<T>() : Sc<T>::zero();
    if (i >= j && i < nb && j < nb) {
      const int ti = t_lower ? i : j, tj = t_lower ? j : i;            // element of T
      const int r = (op == OP_N) ? ti : tj, c = (op == OP_N) ? tj : ti;   // element of A
      const bool referenced = (r == c) ? !unit : (uplo == UPLO_UPPER ? r < c : r > c);
      if (referenced) { v = Ab[r + c * lda]; if (op == OP_C) v = Sc<T>::conj(v); }
      else if (r != c) v = Sc<T>::zero();
    }
    if (i < H && j < H) L11[i * H + j] = (i >= j) ? v : Sc<T>::zero();
    else if (i >= H && j >= H) L22[(i - H) * H + (j - H)] = (i >= j) ? v : Sc<T>::zero();
    else if (i >= H && j < H) L21[(i - H) * H + j] = v;
  }
  __syncthreads();
  if (tid < 2 * H) {
    const T* L = tid < H ? L11 : L22;
    T* X = tid < H ? X11 : X22;
    const int j = tid % H;
    T x[H];
#pragma unroll
    for (int i = 0; i < H; ++i) {
      T acc = (i == j) ? sc_one<T>() : Sc<T>::zero();
#pragma unroll
      for (int q = 0; q < i; ++q) sc_fnma<T>(acc, L[i * H + q], x[q]);   // x[q] = 0 for q < j: uniform control flow
      x[i] = (i < j) ? Sc<T>::zero() : Sc<T>::mul(acc, sc_recip<T>(L[i * H + i]));
    }
#pragma unroll
    for (int i = 0; i < H; ++i) X[i * H + j] = x[i];
  }
  __syncthreads();
  T* M = L11;   // L11 is dead: M = L21 * X11
  for (int idx = tid; idx < H * H; idx += 256) {
    const int i = idx / H, j = idx % H;
    T acc = Sc<T>::zero();
    for (int q = j; q < H; ++q) Sc<T>::fma(acc, L21[i * H + q], X11[q * H + j]);   // X11 is lower: rows q >= j
    M[i * H + j] = acc;
  }
  __syncthreads();
  for (int idx = tid; idx < IB * IB; idx += 256) {
    const int i = idx % IB, j = idx / IB;   // element (i, j) of X = inv(Lc)
    T v = Sc<T>::zero();
    if (i >= j) {
      if (i < H) v = X11[i * H + j];
      else if (j >= H) v = X22[(i - H) * H + (j - H)];
      else {
        T acc = Sc<T>::zero();
        for (int q = 0; q <= i - H; ++q) sc_fnma<T>(acc, X22[(i - H) * H + q], M[q * H + j]);   // X22 lower: columns q <= i - H
        v = acc;
      }
    }
    if (t_lower) Vb[i + j * IB] = v; else Vb[j + i * IB] = v;   // V = X or X^T; the other triangle gets its zeros from (j, i)
    if (i > j) { if (t_lower) Vb[j + i * IB] = Sc<T>::zero(); else Vb[i + j * IB] = Sc<T>::zero(); }
  }
}

// ---- B := alpha * B (alpha == 0: B := 0 without reading it) over an m x n window -----------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) window_scale_kernel(int64_t m, int64_t n, T alpha, bool zero, T* __restrict__ B, int64_t ldb) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= m) return;
  for (int64_t j = blockIdx.y; j < n; j += gridDim.y) {
    T* p = B + i + j * ldb;
    *p = zero ? Sc<T>::zero() : Sc<T>::mul(alpha, *p);
  }
}

// ---- dense image of a symmetric / Hermitian matrix given by one triangle -------------------------------------------------------
// 32 x 32 tiles; a tile in the unreferenced triangle is the (conjugate) transpose of its mirror tile, read coalesced
// and transposed through shared memory.  Hermitian: the imaginary part of the diagonal is taken as zero (BLAS contract).
template <typename T>
__global__ void __launch_bounds__(256) symm_expand_kernel(int uplo, int herm, int64_t n, const T* __restrict__ A, int64_t lda,
                                                          T* __restrict__ W, int64_t ldw) {
  __shared__ T tile[32][33];
  const int64_t ti = blockIdx.x, tj = blockIdx.y;   // tile of W: rows ti*32.., columns tj*32..
  const bool mirrored = (uplo == UPLO_UPPER) ? (ti > tj) : (ti < tj);
  const int64_t si = mirrored ? tj : ti, sj = mirrored ? ti : tj;   // source tile of A
  const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;
  for (int c = ty; c < 32; c += 8) {
    const int64_t gi = si * 32 + tx, gj = sj * 32 + c;
    const bool referenced = gi < n && gj < n && (gi == gj || (uplo == UPLO_UPPER ? gi < gj : gi > gj));
    tile[c][tx] = referenced ? A[gi + gj * lda] : Sc<T>::zero();   // tile[col][row] of the source tile
  }
  __syncthreads();
  for (int c = ty; c < 32; c += 8) {
    const int64_t gi = ti * 32 + tx, gj = tj * 32 + c;
    if (gi >= n || gj >= n) continue;
    T v;
    if (ti != tj) {
      v = mirrored ? tile[tx][c] : tile[c][tx];
      if (mirrored && herm) v = Sc<T>::conj(v);
    } else {
      const bool referenced = (gi == gj) || (uplo == UPLO_UPPER ? gi < gj : gi > gj);
      v = referenced ? tile[c][tx] : tile[tx][c];
      if (!referenced && herm) v = Sc<T>::conj(v);
      if constexpr (sizeof(T) != sizeof(typename Sc<T>::real)) { if (herm && gi == gj) v.y = 0; }
    }
    W[gi + gj * ldw] = v;
  }
}

// sub-vector length held in registers, and sub-vectors per leaf: leaves have order <= NB * NSUB (128; 64 for complex double)
template <typename T> struct LeafOrder { static constexpr int NB = 32; static constexpr int NSUB = 4; };
template <> struct LeafOrder<double2> { static constexpr int NB = 16; static constexpr int NSUB = 4; };

// right-hand sides per CTA of tri_block_solve_kernel: S (LB x (LB+1)) + X (LB x (NV+1)) must fit 227 KB of shared memory
template <typename T> struct BlockSolve { static constexpr int NV = 64; };          // double / complex<float>: 132 + 66.5 KB
template <> struct BlockSolve<float> { static constexpr int NV = 64; };
template <> struct BlockSolve<double2> { static constexpr int NV = 64; };           // LB = 64: 66.5 + 66.5 KB

template <typename T>
T scalar_of(const double a[2]) {
  if constexpr (sizeof(T) == sizeof(typename Sc<T>::real)) return (T)a[0];
  else { T r; r.x = (typename Sc<T>::real)a[0]; r.y = (typename Sc<T>::real)a[1]; return r; }
}

template <typename T, bool SOLVE>
int launch_leaf(const TriProblem& p, bool s_lower, int64_t d0, int nb, cudaStream_t s) {
  constexpr int NB = LeafOrder<T>::NB, NSUB = LeafOrder<T>::NSUB, LB = NB * NSUB;
  const T* A = (const T*)p.A + d0 + d0 * p.lda;
  T* B = p.left ? (T*)p.B + d0 : (T*)p.B + d0 * p.ldb;
  const int64_t nrhs = p.left ? p.n : p.m;
  if (SOLVE && p.Vinv) {   // X_b = inv(T_bb) * B_b (or B_b * inv(T_bb)) out of place on the tensor pipe, then copied back
    const T* Vb = (const T*)p.Vinv + (d0 / LB) * (int64_t)LB * LB;
    GemmProblem g;
    g.type = p.type; g.opa = OP_N; g.opb = OP_N; g.k = nb;
    g.alpha[0] = 1.0; g.alpha[1] = 0.0; g.beta[0] = 0.0; g.beta[1] = 0.0;
    if (p.left) { g.m = nb; g.n = nrhs; g.A = Vb; g.lda = LB; g.B = B; g.ldb = p.ldb; g.C = p.Xtmp; g.ldc = LB; }
    else { g.m = nrhs; g.n = nb; g.A = B; g.lda = p.ldb; g.B = Vb; g.ldb = LB; g.C = p.Xtmp; g.ldc = nrhs; }
    B200_CUDA_TRY(run_gemm_device(g, s, B200BLAS_AUTO));
    return (int)cudaMemcpy2DAsync(B, (size_t)p.ldb * sizeof(T), p.Xtmp, (size_t)g.ldc * sizeof(T), (size_t)g.m * sizeof(T), (size_t)g.n,
                                  cudaMemcpyDeviceToDevice, s);
  }
  if constexpr (SOLVE) {
    static const bool block_solve = [] { const char* e = getenv("B200BLAS_TRSM_LEAF"); return !(e && e[0] == 'v'); }();   // "vector": the kernel above
    if (s_lower && block_solve) {
      constexpr int NV = BlockSolve<T>::NV;
      constexpr size_t bsmem = ((size_t)LB * (LB + 1) + (size_t)LB * (NV + 1)) * sizeof(T);
      B200_SET_MAX_DYN_SMEM_ONCE((tri_block_solve_kernel<T, LB, NV>), bsmem);
      tri_block_solve_kernel<T, LB, NV><<<(unsigned)((nrhs + NV - 1) / NV), 256, bsmem, s>>>(p.left, p.op, p.uplo, p.unit, nb, nrhs, A, p.lda, B, p.ldb);
      count_launch();
      return (int)cudaGetLastError();
    }
  }
  const unsigned grid = (unsigned)((nrhs + LEAF_THREADS - 1) / LEAF_THREADS);
  constexpr size_t smem = (size_t)LB * LB * sizeof(T);
  if (s_lower) {
    B200_SET_MAX_DYN_SMEM_ONCE((tri_leaf_kernel<T, NB, NSUB, true, SOLVE>), smem);
    tri_leaf_kernel<T, NB, NSUB, true, SOLVE><<<grid, LEAF_THREADS, smem, s>>>(p.left, p.op, p.uplo, p.unit, nb, nrhs, A, p.lda, B, p.ldb);
  } else {
    B200_SET_MAX_DYN_SMEM_ONCE((tri_leaf_kernel<T, NB, NSUB, false, SOLVE>), smem);
    tri_leaf_kernel<T, NB, NSUB, false, SOLVE><<<grid, LEAF_THREADS, smem, s>>>(p.left, p.op, p.uplo, p.unit, nb, nrhs, A, p.lda, B, p.ldb);
  }
  count_launch();
  return (int)cudaGetLastError();
}

// block (r0.., c0..) of T = op(A) as a GEMM operand
template <typename T>
const void* tblock(const TriProblem& p, int64_t r0, int64_t c0) {
  return (p.op == OP_N) ? (const void*)((const T*)p.A + r0 + c0 * p.lda) : (const void*)((const T*)p.A + c0 + r0 * p.lda);
}

// dst(rows rd.., or columns) += sign * T(block) * src   /   src * T(block)
template <typename T>
int offdiag_update(const TriProblem& p, int64_t dst0, int64_t ndst, int64_t src0, int64_t nsrc, double sign, cudaStream_t s) {
  GemmProblem g;
  g.type = p.type;
  g.alpha[0] = sign; g.alpha[1] = 0.0; g.beta[0] = 1.0; g.beta[1] = 0.0;
  if (p.left) {   // B[dst rows, :] += sign * T[dst, src] * B[src rows, :]
    g.opa = p.op; g.opb = OP_N; g.m = ndst; g.n = p.n; g.k = nsrc;
    g.A = tblock<T>(p, dst0, src0); g.lda = p.lda;
    g.B = (const T*)p.B + src0; g.ldb = p.ldb;
    g.C = (T*)p.B + dst0; g.ldc = p.ldb;
  } else {        // B[:, dst cols] += sign * B[:, src cols] * T[src, dst]
    g.opa = OP_N; g.opb = p.op; g.m = p.m; g.n = ndst; g.k = nsrc;
    g.A = (const T*)p.B + src0 * p.ldb; g.lda = p.ldb;
    g.B = tblock<T>(p, src0, dst0); g.ldb = p.lda;
    g.C = (T*)p.B + dst0 * p.ldb; g.ldc = p.ldb;
  }
  return run_gemm_device(g, s, B200BLAS_AUTO);
}

static int64_t split_point(int64_t d, int nb) {   // largest nb * 2^j strictly below d
  int64_t h = nb;
  while (h * 2 < d) h *= 2;
  return h;
}

// Triangle rows/columns [d0, d0 + d).  t_lower: op(A) is lower triangular.
template <typename T, bool SOLVE>
int tri_recurse(const TriProblem& p, bool t_lower, int64_t d0, int64_t d, cudaStream_t s) {
  constexpr int LB = LeafOrder<T>::NB * LeafOrder<T>::NSUB;
  const bool s_lower = p.left ? t_lower : !t_lower;   // S = T (left) or T^T (right)
  if (d <= LB) return launch_leaf<T, SOLVE>(p, s_lower, d0, (int)d, s);
  const int64_t d1 = split_point(d, LB), d2 = d - d1;
  // which diagonal block goes first: in a solve, the one whose unknowns feed the other; in a product (in place), the
  // one whose inputs are not needed by the other any more
  // solve, left:  T lower -> (1) first; T upper -> (2) first.   solve, right: X T = B: T lower -> (2) first; upper -> (1).
  // multiply, left:  B2' = T21 B1 + T22 B2 (lower) -> (2) first; upper -> (1) first.   right: lower -> (1) first; upper -> (2).
  const bool first_is_1 = SOLVE ? (p.left ? t_lower : !t_lower) : (p.left ? !t_lower : t_lower);
  const int64_t f0 = first_is_1 ? d0 : d0 + d1, fd = first_is_1 ? d1 : d2;
  const int64_t g0 = first_is_1 ? d0 + d1 : d0, gd = first_is_1 ? d2 : d1;
  if constexpr (SOLVE) {
    B200_CUDA_TRY((tri_recurse<T, SOLVE>(p, t_lower, f0, fd, s)));
    B200_CUDA_TRY(offdiag_update<T>(p, g0, gd, f0, fd, -1.0, s));      // remaining block -= T[.,.] * solved block
    B200_CUDA_TRY((tri_recurse<T, SOLVE>(p, t_lower, g0, gd, s)));
  } else {
    B200_CUDA_TRY((tri_recurse<T, SOLVE>(p, t_lower, f0, fd, s)));      // diagonal part of the first block
    B200_CUDA_TRY(offdiag_update<T>(p, f0, fd, g0, gd, 1.0, s));       // first block += T[.,.] * (still original) other block
    B200_CUDA_TRY((tri_recurse<T, SOLVE>(p, t_lower, g0, gd, s)));
  }
  return 0;
}

template <typename T>
int scale_window(int64_t m, int64_t n, const double alpha[2], T* B, int64_t ldb, cudaStream_t s) {
  if (alpha[0] == 1.0 && alpha[1] == 0.0) return 0;
  const bool zero = alpha[0] == 0.0 && alpha[1] == 0.0;
  dim3 grid((unsigned)((m + 255) / 256), (unsigned)std::min<int64_t>(n, 1024));
  window_scale_kernel<T><<<grid, 256, 0, s>>>(m, n, scalar_of<T>(alpha), zero, B, ldb);
  count_launch();
  return (int)cudaGetLastError();
}

template <typename T, bool SOLVE>
int run_tri(const TriProblem& p, cudaStream_t s) {
  const bool zero = p.alpha[0] == 0.0 && p.alpha[1] == 0.0;
  if (!zero) {   // alpha == 0: the result is zero whatever A holds (netlib ?TRSM/?TRMM quick path)
    const bool t_lower = (p.uplo == UPLO_LOWER) == (p.op == OP_N);
    // leaves: B200BLAS_TRSM=inv forces the inverse-based leaves (one batched inversion of all diagonal blocks, then every
    // leaf is a product on the tensor-pipe kernels), =subst the substitution leaf; default: inverse leaves for large solves
    // (measured on B200, profiles/bench_r02: dtrsm 8192 21.9 -> 28.3 TFLOP/s at best), substitution where the inversion of the
    // whole triangle would not be amortised (the panel solves inside ?potrf_ / ?getrf_)
    static const int leaf_env = [] { const char* e = getenv("B200BLAS_TRSM"); return !e ? 0 : (e[0] == 'i' ? 1 : (e[0] == 's' ? 2 : 0)); }();
    constexpr int LB = LeafOrder<T>::NB * LeafOrder<T>::NSUB;
    const int64_t na = p.left ? p.m : p.n, nrhs = p.left ? p.n : p.m;
    // (final pass of round 2: with the shared-memory block solve as the leaf, a lower canonical solve runs at 26.6 TFLOP/s
    // without any workspace, run to run; the inverse path measured 19 - 28 TFLOP/s depending on the stream-ordered
    // allocator, so it is kept for the UPPER canonical solves only, whose leaf is still the one-vector-per-thread kernel)
    const bool s_lower_all = p.left ? t_lower : !t_lower;
    const bool use_inv = leaf_env == 1 || (leaf_env == 0 && !s_lower_all && na >= 2048 && nrhs >= 1024);
    if (SOLVE && p.Vinv && p.Xtmp) {   // the caller (lapack.cu) already holds the inverses of the diagonal blocks
      B200_CUDA_TRY((tri_recurse<T, SOLVE>(p, t_lower, 0, na, s)));
    } else if (SOLVE && use_inv) {
      const int64_t nblocks = (na + LB - 1) / LB;
      const size_t vbytes = (size_t)nblocks * LB * LB * sizeof(T), xbytes = (size_t)LB * (size_t)nrhs * sizeof(T);
      unsigned char* ws = nullptr;
      B200_CUDA_TRY(cudaMallocAsync((void**)&ws, vbytes + xbytes + 256, s));
      constexpr size_t smem = 5 * (size_t)(LB / 2) * (LB / 2) * sizeof(T);
      B200_SET_MAX_DYN_SMEM_ONCE((trtri_diag_kernel<T, LB>), smem);
      trtri_diag_kernel<T, LB><<<(unsigned)nblocks, 256, smem, s>>>(p.op, p.uplo, p.unit, t_lower ? 1 : 0, na, (const T*)p.A, p.lda, (T*)ws);
      count_launch();
      int e = (int)cudaGetLastError();
      TriProblem q = p;
      q.Vinv = ws; q.Xtmp = ws + (vbytes + 255) / 256 * 256;
      if (!e) e = tri_recurse<T, SOLVE>(q, t_lower, 0, na, s);
      cudaFreeAsync(ws, s);
      if (e) return e;
    } else {
      B200_CUDA_TRY((tri_recurse<T, SOLVE>(p, t_lower, 0, p.left ? p.m : p.n, s)));
    }
  }
  // the reference scales after the solve / inside the product (blas/level3_impl.h:174-175, :278-281)
  return scale_window<T>(p.m, p.n, p.alpha, (T*)p.B, p.ldb, s);
}

template <bool SOLVE>
int dispatch_tri(const TriProblem& p, cudaStream_t s) {
  if (p.m <= 0 || p.n <= 0) return 0;
  note_variant(SOLVE ? "trsm_recursive_leaf+gemm" : "trmm_recursive_leaf+gemm");
  switch (p.type) {
    case TY_S: return run_tri<float, SOLVE>(p, s);
    case TY_D: return run_tri<double, SOLVE>(p, s);
    case TY_C: return run_tri<float2, SOLVE>(p, s);
    default: return run_tri<double2, SOLVE>(p, s);
  }
}

template <typename T>
int expand_typed(const SymmProblem& p, int64_t na, void* W, int64_t ldw, cudaStream_t s) {
  const unsigned t = (unsigned)((na + 31) / 32);
  symm_expand_kernel<T><<<dim3(t, t), 256, 0, s>>>(p.uplo, p.herm, na, (const T*)p.A, p.lda, (T*)W, ldw);
  count_launch();
  return (int)cudaGetLastError();
}

}  // namespace

int launch_trsm(const TriProblem& p, cudaStream_t s) { return dispatch_tri<true>(p, s); }

// inverses of the diagonal leaf blocks of op(A) (order trsm_leaf_order(type) each, stored densely one after the other in V):
// what TriProblem::Vinv expects.  Lets ?getrf_ invert a panel's L11 once and use it for several solves on different streams.
int trsm_leaf_order(int type) {
  return type == TY_Z ? LeafOrder<double2>::NB * LeafOrder<double2>::NSUB : LeafOrder<double>::NB * LeafOrder<double>::NSUB;
}
template <typename T>
static int trtri_typed(const TriProblem& p, void* V, cudaStream_t s) {
  constexpr int LB = LeafOrder<T>::NB * LeafOrder<T>::NSUB;
  const bool t_lower = (p.uplo == UPLO_LOWER) == (p.op == OP_N);
  const int64_t na = p.left ? p.m : p.n;
  constexpr size_t smem = 5 * (size_t)(LB / 2) * (LB / 2) * sizeof(T);
  B200_SET_MAX_DYN_SMEM_ONCE((trtri_diag_kernel<T, LB>), smem);
  trtri_diag_kernel<T, LB><<<(unsigned)((na + LB - 1) / LB), 256, smem, s>>>(p.op, p.uplo, p.unit, t_lower ? 1 : 0, na, (const T*)p.A, p.lda, (T*)V);
  count_launch();
  return (int)cudaGetLastError();
}
int launch_trtri_diag(const TriProblem& p, void* V, cudaStream_t s) {
  switch (p.type) {
    case TY_S: return trtri_typed<float>(p, V, s);
    case TY_D: return trtri_typed<double>(p, V, s);
    case TY_C: return trtri_typed<float2>(p, V, s);
    default: return trtri_typed<double2>(p, V, s);
  }
}
bool trsm_substitution_forced() {
  static const bool v = [] { const char* e = getenv("B200BLAS_TRSM"); return e && e[0] == 's'; }();
  return v;
}
bool trsm_inverse_forced() {
  static const bool v = [] { const char* e = getenv("B200BLAS_TRSM"); return e && e[0] == 'i'; }();
  return v;
}
int launch_trmm(const TriProblem& p, cudaStream_t s) { return dispatch_tri<false>(p, s); }

size_t symm_workspace_bytes(const SymmProblem& p) {
  const int64_t na = p.left ? p.m : p.n;
  const int64_t q = 32 / type_bytes(p.type) > 0 ? 32 / type_bytes(p.type) : 1;
  const int64_t ldw = (na + q - 1) / q * q;
  return (size_t)ldw * (size_t)na * (size_t)type_bytes(p.type);
}

// C = alpha * A * B + beta * C (left) or alpha * B * A + beta * C (right); A symmetric / Hermitian, one triangle stored
int launch_symm(const SymmProblem& p, cudaStream_t s, void* workspace) {
  if (p.m <= 0 || p.n <= 0) return 0;
  const int64_t na = p.left ? p.m : p.n;
  const int64_t q = 32 / type_bytes(p.type) > 0 ? 32 / type_bytes(p.type) : 1;
  const int64_t ldw = (na + q - 1) / q * q;
  int e;
  switch (p.type) {
    case TY_S: e = expand_typed<float>(p, na, workspace, ldw, s); break;
    case TY_D: e = expand_typed<double>(p, na, workspace, ldw, s); break;
    case TY_C: e = expand_typed<float2>(p, na, workspace, ldw, s); break;
    default: e = expand_typed<double2>(p, na, workspace, ldw, s); break;
  }
  if (e) return e;
  GemmProblem g;
  g.type = p.type; g.opa = OP_N; g.opb = OP_N; g.m = p.m; g.n = p.n; g.k = na;
  g.alpha[0] = p.alpha[0]; g.alpha[1] = p.alpha[1]; g.beta[0] = p.beta[0]; g.beta[1] = p.beta[1];
  if (p.left) { g.A = workspace; g.lda = ldw; g.B = p.B; g.ldb = p.ldb; }
  else { g.A = p.B; g.lda = p.ldb; g.B = workspace; g.ldb = ldw; }
  g.C = p.C; g.ldc = p.ldc;
  return run_gemm_device(g, s, B200BLAS_AUTO);
}

}  // namespace b200
