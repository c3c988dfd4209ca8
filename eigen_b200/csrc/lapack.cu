// eigen_b200/csrc/lapack.cu -- device-resident blocked Cholesky and LU on the GEMM kernels (sm_100a).
//
// SURVEY.md section 8 row f3.  Reference being replaced:
//   llt_inplace<Scalar,UpLo>::blocked / unblocked     Eigen/src/Cholesky/LLT.h:299-360      (?potrf_, lapack/cholesky.cpp:14-38)
//   partial_lu_impl::blocked_lu / unblocked_lu        Eigen/src/LU/PartialPivLU.h:361-496   (?getrf_, lapack/lu.cpp:14-42)
// The reference walks the diagonal in blocks of <= 128 / 256 columns: unblocked factor of the diagonal block, a
// triangular solve against it, and a rank-k update of the trailing matrix ("bottleneck", LLT.h:357; PartialPivLU.h:492).
// With the whole matrix resident in HBM the same three steps are arranged recursively so that the solve and the update
// are as large as possible (half of the current range each time) and run on the tensor-pipe kernels through
// launch_trsm / run_gemm_device; only the leaves are special kernels:
//   potf2_leaf_kernel    Cholesky of a diagonal block of order <= 32, one warp, block in shared memory;
//   getf2_panel_kernel   partial-pivoting LU of an (m - j0) x (<= 32) panel by ONE cooperative grid: every CTA keeps its
//                        slab of panel rows in shared memory, and each column costs one grid-wide barrier -- the CTAs
//                        publish their best pivot candidate together with that candidate's whole row, so after the
//                        barrier every CTA knows the pivot row's contents and updates its slab without further traffic.
// Row interchanges outside the panel (xLASWP) are turned into a permutation first (sequential simulation in shared
// memory) and applied as parallel gathers, instead of a chain of dependent row swaps.
#include <climits>
#include <cooperative_groups.h>
#include <cstdlib>

#include "../../include/b200blas.h"
#include "common.cuh"
#include "scalar.cuh"

namespace b200 {
namespace {

template <typename T> struct Leaf { static constexpr int NB = 32; };
template <> struct Leaf<double2> { static constexpr int NB = 16; };

template <typename T> __device__ __forceinline__ typename Sc<T>::real sc_real(T a) {
  if constexpr (sizeof(T) == sizeof(typename Sc<T>::real)) return a; else return a.x;
}
template <typename T> __device__ __forceinline__ T sc_from_real(typename Sc<T>::real r) {
  if constexpr (sizeof(T) == sizeof(typename Sc<T>::real)) return r; else { T v; v.x = r; v.y = 0; return v; }
}
template <typename T> __device__ __forceinline__ T sc_scale(T a, typename Sc<T>::real r) {
  if constexpr (sizeof(T) == sizeof(typename Sc<T>::real)) return a * r; else { T v; v.x = a.x * r; v.y = a.y * r; return v; }
}
template <typename T> __device__ __forceinline__ T sc_div_real(T a, typename Sc<T>::real r) {
  if constexpr (sizeof(T) == sizeof(typename Sc<T>::real)) return a / r; else { T v; v.x = a.x / r; v.y = a.y / r; return v; }
}
// |a| as the pivot score (scalar_score_coeff_op = abs, Eigen/src/Core/functors/UnaryFunctors.h)
template <typename T> __device__ __forceinline__ double sc_score(T a) {
  if constexpr (sizeof(T) == sizeof(typename Sc<T>::real)) return fabs((double)a);
  else return hypot((double)a.x, (double)a.y);
}
template <typename T> __device__ __forceinline__ T sc_div(T a, T b) {
  if constexpr (sizeof(T) == sizeof(typename Sc<T>::real)) return a / b;
  else return Sc<T>::mul(a, sc_recip<T>(b));
}

// ---- Cholesky leaf -------------------------------------------------------------------------------------------------------
// Right-looking Cholesky of a d x d diagonal block (d <= NB) by one warp: lane i keeps row i of the lower-canonical
// block in registers (a[j] = element (i, j), j <= i; element (i, j) = A(i,j) for uplo = L, conj(A(j,i)) for uplo = U --
// the reference factors the transpose for Upper, LLT.h:367-380), columns travel between lanes by shuffles.
// A non-positive pivot at column k stores d0 + k + 1 into *info (smallest wins) and stops, like LLT.h:316-317.
template <typename T> __device__ __forceinline__ T sc_shfl(T v, int src) {
  if constexpr (sizeof(T) == sizeof(typename Sc<T>::real)) return __shfl_sync(0xffffffffu, v, src);
  else { T r; r.x = __shfl_sync(0xffffffffu, v.x, src); r.y = __shfl_sync(0xffffffffu, v.y, src); return r; }
}

// sqrt(x) and 1/sqrt(x) for the pivot.  The IEEE double sqrt and divide are multi-hundred-cycle software sequences on
// the critical path of every column; a float rsqrt seed refined by two Newton steps in double (relative error ~1e-16,
// then one correction of the root) costs a few dependent FMAs.  Out-of-float-range pivots take the IEEE path.
__device__ __forceinline__ void pivot_roots(double x, double& root, double& inv_root) {
  if (x > 1e-30 && x < 1e30) {
    double r = (double)rsqrtf((float)x);
    r = r * fma(-0.5 * x, r * r, 1.5);
    r = r * fma(-0.5 * x, r * r, 1.5);
    double l = x * r;
    l = fma(0.5 * r, fma(-l, l, x), l);
    root = l; inv_root = r;
  } else {
    root = sqrt(x); inv_root = 1.0 / root;
  }
}
__device__ __forceinline__ void pivot_roots(float x, float& root, float& inv_root) { root = sqrtf(x); inv_root = 1.0f / root; }

template <typename T, int NB>
__global__ void __launch_bounds__(32) potf2_leaf_kernel(int upper, int d, int64_t d0, T* __restrict__ A, int64_t lda, int* __restrict__ info) {
  using R = typename Sc<T>::real;
  const int lane = threadIdx.x;
  T a[NB];   // the lane's whole row in flight at once
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    a[j] = (j < d && lane < d && lane >= j) ? (upper ? A[j + (int64_t)lane * lda] : A[lane + (int64_t)j * lda]) : Sc<T>::zero();
    if (upper) a[j] = Sc<T>::conj(a[j]);
  }
  bool ok = true;   // warp-uniform
#pragma unroll
  for (int k = 0; k < NB; ++k) {
    if (k < d && ok) {
      const R x = sc_real<T>(sc_shfl<T>(a[k], k));
      if (x <= (R)0) {   // a NaN pivot continues, as in the reference
        if (lane == 0) atomicMin(info, (int)(d0 + k + 1));
        ok = false;
      } else {
        R l, rl;
        pivot_roots(x, l, rl);
        if (lane == k) a[k] = sc_from_real<T>(l);
        else if (lane > k) a[k] = sc_scale<T>(a[k], rl);
#pragma unroll
        for (int j = k + 1; j < NB; ++j) {
          const T ljk = sc_shfl<T>(a[k], j);   // L(j, k), final
          if (lane >= j) sc_fnma<T>(a[j], a[k], Sc<T>::conj(ljk));
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < NB; ++j)
    if (j < d && lane < d && lane >= j) {
      if (upper) A[j + (int64_t)lane * lda] = Sc<T>::conj(a[j]); else A[lane + (int64_t)j * lda] = a[j];
    }
}

// ---- DRAFT (round 2, not yet run on hardware; opt-in with B200BLAS_POTF2=cta) --------------------------------------------
// CTA-wide Cholesky leaf of order d <= NBL: the block lives in shared memory, every column costs two __syncthreads and
// its trailing update is spread over all 256 threads (the one-warp leaf issues ~8k dependent instructions per 32 x 32
// block = 34 us; profiles/launches_r01_dpotrf8192_v2.md).
template <typename T, int NBL>
__global__ void __launch_bounds__(256) potf2_cta_kernel(int upper, int d, int64_t d0, T* __restrict__ A, int64_t lda, int* __restrict__ info) {
  using R = typename Sc<T>::real;
  __shared__ T S[NBL][NBL + 1];   // lower-canonical block, S[i][j] for i >= j
  __shared__ int failed;
  const int tid = threadIdx.x;
  if (tid == 0) failed = 0;
  for (int idx = tid; idx < NBL * NBL; idx += 256) {
    const int i = idx % NBL, j = idx / NBL;
    if (i < d && j <= i) {
      const T v = upper ? Sc<T>::conj(A[j + (int64_t)i * lda]) : A[i + (int64_t)j * lda];
      S[i][j] = v;
    }
  }
  __syncthreads();
  const int ti = tid % NBL, tg = tid / NBL;          // row owned in the trailing update, column group
  constexpr int GROUPS = 256 / NBL;
  for (int k = 0; k < d; ++k) {
    const R x = sc_real<T>(S[k][k]);
    if (x <= (R)0) {   // uniform: every thread reads the same value
      if (tid == 0) { atomicMin(info, (int)(d0 + k + 1)); failed = 1; }
      break;
    }
    R l, rl;
    pivot_roots(x, l, rl);
    __syncthreads();                                  // everybody has read S[k][k]
    if (tid == k) S[k][k] = sc_from_real<T>(l);
    else if (tid > k && tid < d) S[tid][k] = sc_scale<T>(S[tid][k], rl);
    __syncthreads();                                  // column k is final
    if (ti < d) {
      const T lik = S[ti][k];
      for (int j = k + 1 + tg; j <= ti; j += GROUPS) sc_fnma<T>(S[ti][j], lik, Sc<T>::conj(S[j][k]));
    }
    __syncthreads();
  }
  __syncthreads();
  for (int idx = tid; idx < NBL * NBL; idx += 256) {
    const int i = idx % NBL, j = idx / NBL;
    if (i < d && j <= i) {
      if (upper) A[j + (int64_t)i * lda] = Sc<T>::conj(S[i][j]); else A[i + (int64_t)j * lda] = S[i][j];
    }
  }
}

static int64_t split_point(int64_t d, int nb) {
  int64_t h = nb;
  while (h * 2 < d) h *= 2;
  return h;
}

template <typename T>
int potrf_rec(const PotrfProblem& p, int64_t d0, int64_t d, cudaStream_t s) {
  constexpr int NB = Leaf<T>::NB;
  const bool upper = p.uplo == UPLO_UPPER;
  const bool cplx = sizeof(T) != sizeof(typename Sc<T>::real);
  T* A = (T*)p.A;
  static const bool cta_leaf = [] { const char* e = getenv("B200BLAS_POTF2"); return e && e[0] == 'c'; }();   // DRAFT, opt-in
  constexpr int NBL = 64 / (sizeof(T) == 16 ? 2 : 1);   // 64 x 65 elements of <= 8 bytes, 32 x 33 of 16 bytes
  if (cta_leaf && d <= NBL) {
    potf2_cta_kernel<T, NBL><<<1, 256, 0, s>>>(upper ? 1 : 0, (int)d, d0, A + d0 + d0 * p.lda, p.lda, p.dinfo);
    count_launch();
    return (int)cudaGetLastError();
  }
  if (d <= NB) {
    potf2_leaf_kernel<T, NB><<<1, 32, 0, s>>>(upper ? 1 : 0, (int)d, d0, A + d0 + d0 * p.lda, p.lda, p.dinfo);
    count_launch();
    return (int)cudaGetLastError();
  }
  const int64_t d1 = split_point(d, cta_leaf ? NBL : NB), d2 = d - d1;
  B200_CUDA_TRY(potrf_rec<T>(p, d0, d1, s));
  TriProblem t;
  t.type = p.type; t.uplo = p.uplo; t.op = OP_C; t.unit = 0; t.alpha[0] = 1.0; t.alpha[1] = 0.0;
  t.A = A + d0 + d0 * p.lda; t.lda = p.lda; t.ldb = p.lda;
  GemmProblem g;
  g.type = p.type; g.alpha[0] = -1.0; g.alpha[1] = 0.0; g.beta[0] = 1.0; g.beta[1] = 0.0;
  g.m = d2; g.n = d2; g.k = d1; g.uplo = p.uplo; g.herm = cplx ? 1 : 0;
  g.C = A + (d0 + d1) + (d0 + d1) * p.lda; g.ldc = p.lda; g.lda = p.lda; g.ldb = p.lda;
  if (!upper) {   // A21 := A21 * L11^-H ; A22 -= A21 * A21^H   (LLT.h:356-357)
    t.left = 0; t.m = d2; t.n = d1; t.B = A + (d0 + d1) + d0 * p.lda;
    g.opa = OP_N; g.opb = OP_C; g.A = t.B; g.B = t.B;
  } else {        // A12 := U11^-H * A12 ; A22 -= A12^H * A12
    t.left = 1; t.m = d1; t.n = d2; t.B = A + d0 + (d0 + d1) * p.lda;
    g.opa = OP_C; g.opb = OP_N; g.A = t.B; g.B = t.B;
  }
  B200_CUDA_TRY(launch_trsm(t, s));
  B200_CUDA_TRY(run_gemm_device(g, s, B200BLAS_AUTO));
  return potrf_rec<T>(p, d0 + d1, d2, s);
}

// ---- LU panel ------------------------------------------------------------------------------------------------------------
// Scratch layout (device memory, per factorization):
//   unsigned barrier counter | per parity (2): { per CTA: double score, int row, T row[NBP] } + T rowk[NBP]
template <typename T, int NBP> struct PanelCand { double score; int row; int pad; T vals[NBP]; };

__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned v;
    do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory"); } while (v < target);
    __threadfence();
  }
  __syncthreads();
}

// Panel: rows [0, mrows) x columns [0, nb) at A (panel-relative).  ipiv receives 1-based GLOBAL row numbers
// (row_base + pivot + 1).  info: smallest col_base + k + 1 with an exactly zero pivot column (PartialPivLU.h:396-401).
template <typename T, int NBP>
__global__ void __launch_bounds__(256)
getf2_panel_kernel(int64_t mrows, int nb, int rows_per_cta, T* __restrict__ A, int64_t lda, int* __restrict__ ipiv, int64_t row_base,
                   int* __restrict__ info, int64_t col_base, unsigned char* __restrict__ scratch, T* __restrict__ gslab) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int LDS = NBP + 1;
  // slab[r * (NBP + 1) + c]: shared memory, or -- panels too tall for G slabs of shared memory (m above ~120k rows
  // for double) -- this CTA's stretch of a global scratch buffer (L2-resident; slower, but no row limit)
  T* slab = gslab ? gslab + (size_t)blockIdx.x * (size_t)rows_per_cta * LDS : reinterpret_cast<T*>(smem_raw);
  __shared__ T prow[NBP], krow[NBP];
  __shared__ double red_score[8];
  __shared__ int red_row[8], red_w[8];
  __shared__ double win_score;
  __shared__ int win_row;
  using Cand = PanelCand<T, NBP>;
  unsigned* counter = reinterpret_cast<unsigned*>(scratch);
  const int G = gridDim.x, cta = blockIdx.x, tid = threadIdx.x;
  Cand* cands[2];
  T* rowk[2];
  {
    unsigned char* base = scratch + 64;
    const size_t per = sizeof(Cand) * (size_t)G + sizeof(T) * NBP;
    for (int q = 0; q < 2; ++q) {
      cands[q] = reinterpret_cast<Cand*>(base + q * ((per + 63) / 64 * 64));
      rowk[q] = reinterpret_cast<T*>(reinterpret_cast<unsigned char*>(cands[q]) + sizeof(Cand) * (size_t)G);
    }
  }
  const int64_t r0 = (int64_t)cta * rows_per_cta;
  const int R = (int)max((int64_t)0, min((int64_t)rows_per_cta, mrows - r0));
  for (int r = tid; r < R; r += 256)
    for (int c0 = 0; c0 < nb; c0 += 8) {
      T v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = (c0 + q < nb) ? A[(r0 + r) + (int64_t)(c0 + q) * lda] : Sc<T>::zero();
#pragma unroll
      for (int q = 0; q < 8; ++q) if (c0 + q < nb) slab[r * LDS + c0 + q] = v[q];
    }
  __syncthreads();
  const int steps = (int)min((int64_t)nb, mrows);
  for (int k = 0; k < steps; ++k) {
    const int par = k & 1;
    // 1. local pivot candidate among rows >= k: largest |a(r,k)|, smallest row on ties (maxCoeff keeps the first)
    double best = -1.0;
    int brow = INT_MAX;
    for (int r = tid; r < R; r += 256) {
      const int64_t gr = r0 + r;
      if (gr < k) continue;
      const double sc = sc_score<T>(slab[r * LDS + k]);
      if (sc > best || (sc == best && (int)gr < brow)) { best = sc; brow = (int)gr; }
    }
    for (int off = 16; off > 0; off >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, off);
      const int orow = __shfl_xor_sync(0xffffffffu, brow, off);
      if (ob > best || (ob == best && orow < brow)) { best = ob; brow = orow; }
    }
    if ((tid & 31) == 0) { red_score[tid >> 5] = best; red_row[tid >> 5] = brow; }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < 8; ++w)
        if (red_score[w] > best || (red_score[w] == best && red_row[w] < brow)) { best = red_score[w]; brow = red_row[w]; }
      win_score = best; win_row = brow;
    }
    __syncthreads();
    // 2. publish the candidate (score, row, whole row contents) and, from its owner, row k
    {
      const double ls = win_score;
      const int lr = win_row;
      if (tid == 0) { cands[par][cta].score = ls; cands[par][cta].row = lr; }
      if (ls >= 0.0 && tid < nb) cands[par][cta].vals[tid] = slab[(lr - (int)r0) * LDS + tid];
      if (k >= r0 && k < r0 + R && tid < nb) rowk[par][tid] = slab[(k - (int)r0) * LDS + tid];
    }
    // 3. one grid-wide barrier per column
    grid_barrier(counter, (unsigned)(k + 1) * (unsigned)G);
    // 4. every CTA reduces the published candidates (identically): one L2 load per thread, then a block reduction;
    //    the winner's row and row k are fetched by the first warp
    {
      double gb = -1.0;
      int grow = INT_MAX, gw = -1;
      if (tid < G) {
        gb = __ldcg(&cands[par][tid].score);   // L2 reads: the data was written by other SMs
        grow = __ldcg(&cands[par][tid].row);
        gw = tid;
        if (!(gb >= 0.0)) { gb = -1.0; grow = INT_MAX; gw = -1; }
      }
      for (int off = 16; off > 0; off >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, gb, off);
        const int orow = __shfl_xor_sync(0xffffffffu, grow, off);
        const int ow = __shfl_xor_sync(0xffffffffu, gw, off);
        if (ob > gb || (ob == gb && orow < grow)) { gb = ob; grow = orow; gw = ow; }
      }
      if ((tid & 31) == 0) { red_score[tid >> 5] = gb; red_row[tid >> 5] = grow; red_w[tid >> 5] = gw; }
      __syncthreads();
      if (tid < 32) {
        gb = red_score[0]; grow = red_row[0]; gw = red_w[0];
        for (int w = 1; w < 8; ++w)
          if (red_score[w] > gb || (red_score[w] == gb && red_row[w] < grow)) { gb = red_score[w]; grow = red_row[w]; gw = red_w[w]; }
        if (gw < 0) { gb = 0.0; grow = k; }   // a column of NaNs: no candidate compares greater; treated as a zero pivot
        if (tid == 0) { win_score = gb; win_row = grow; }
        if (tid < nb && gw >= 0) { prow[tid] = __ldcg(&cands[par][gw].vals[tid]); krow[tid] = __ldcg(&rowk[par][tid]); }
      }
    }
    __syncthreads();
    const double gscore = win_score;
    const int piv = win_row;
    if (cta == 0 && tid == 0) {
      ipiv[k] = (int)(row_base + piv + 1);
      if (gscore == 0.0) atomicMin(info, (int)(col_base + k + 1));
    }
    if (gscore != 0.0) {
      // 5. interchange rows k and piv inside the panel (PartialPivLU.h:384-388)
      if (piv != k) {
        if (k >= r0 && k < r0 + R && tid < nb) slab[(k - (int)r0) * LDS + tid] = prow[tid];
        if (piv >= r0 && piv < r0 + R && tid < nb) slab[(piv - (int)r0) * LDS + tid] = krow[tid];
      }
      __syncthreads();
      // 6. scale the column below the pivot and update the rest of the panel (PartialPivLU.h:392, :404-405)
      const T pv = prow[k];
      for (int r = tid; r < R; r += 256) {
        if (r0 + r <= k) continue;
        T* row = slab + r * LDS;
        const T l = sc_div<T>(row[k], pv);
        row[k] = l;
        for (int j = k + 1; j < nb; ++j) sc_fnma<T>(row[j], l, prow[j]);
      }
    }
    __syncthreads();
  }
  for (int c = 0; c < nb; ++c)
    for (int r = tid; r < R; r += 256) A[(r0 + r) + (int64_t)c * lda] = slab[r * LDS + c];
}

// ---- DRAFT (round 2, compiled but not yet run on hardware; opt-in with B200BLAS_GETF2=cluster) ---------------------------
// The same panel factorization on ONE thread-block cluster: the slab of every CTA lives in its shared memory, the
// per-column exchange goes through distributed shared memory and the barrier is the hardware cluster barrier instead
// of a global-atomic grid barrier (4.1 us per column in round 1, profiles/launches_r01_dgetrf8192_v2.md).
// Per column: local arg-max -> each CTA writes its candidate (score, row, row contents; the owner of row k also row k)
// into slot [my rank] of EVERY CTA's candidate table (remote stores) -> cluster.sync() -> identical local reduction ->
// swap / scale / update.  Tables are double-buffered by column parity, so one barrier per column suffices.
namespace cg = cooperative_groups;

template <typename T, int NBP, int MAXCL>
struct ClusterTables {
  double score[2][MAXCL];
  int row[2][MAXCL];
  T vals[2][MAXCL][NBP];
  T rowk[2][NBP];
};

template <typename T, int NBP, int MAXCL>
__global__ void __launch_bounds__(256)
getf2_cluster_kernel(int64_t mrows, int nb, int rows_per_cta, T* __restrict__ A, int64_t lda, int* __restrict__ ipiv, int64_t row_base,
                     int* __restrict__ info, int64_t col_base) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* slab = reinterpret_cast<T*>(smem_raw);   // slab[r * (NBP + 1) + c]
  constexpr int LDS = NBP + 1;
  __shared__ ClusterTables<T, NBP, MAXCL> tab;
  __shared__ T prow[NBP], krow[NBP];
  __shared__ double red_score[8];
  __shared__ int red_row[8];
  __shared__ double win_score;
  __shared__ int win_row;
  cg::cluster_group cluster = cg::this_cluster();
  const int CL = (int)cluster.num_blocks(), cta = (int)cluster.block_rank(), tid = threadIdx.x;
  const int64_t r0 = (int64_t)cta * rows_per_cta;
  const int R = (int)max((int64_t)0, min((int64_t)rows_per_cta, mrows - r0));
  for (int r = tid; r < R; r += 256)
    for (int c0 = 0; c0 < nb; c0 += 8) {
      T v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = (c0 + q < nb) ? A[(r0 + r) + (int64_t)(c0 + q) * lda] : Sc<T>::zero();
#pragma unroll
      for (int q = 0; q < 8; ++q) if (c0 + q < nb) slab[r * LDS + c0 + q] = v[q];
    }
  cluster.sync();   // every CTA's shared memory is live before anybody writes into it remotely
  const int steps = (int)min((int64_t)nb, mrows);
  for (int k = 0; k < steps; ++k) {
    const int par = k & 1;
    double best = -1.0;
    int brow = INT_MAX;
    for (int r = tid; r < R; r += 256) {
      const int64_t gr = r0 + r;
      if (gr < k) continue;
      const double sc = sc_score<T>(slab[r * LDS + k]);
      if (sc > best || (sc == best && (int)gr < brow)) { best = sc; brow = (int)gr; }
    }
    for (int off = 16; off > 0; off >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, off);
      const int orow = __shfl_xor_sync(0xffffffffu, brow, off);
      if (ob > best || (ob == best && orow < brow)) { best = ob; brow = orow; }
    }
    if ((tid & 31) == 0) { red_score[tid >> 5] = best; red_row[tid >> 5] = brow; }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < 8; ++w)
        if (red_score[w] > best || (red_score[w] == best && red_row[w] < brow)) { best = red_score[w]; brow = red_row[w]; }
      win_score = best; win_row = brow;
    }
    __syncthreads();
    {
      // publish into every CTA's table (thread t handles peer t % CL, value index t / CL ... simple strided loops)
      const double ls = win_score;
      const int lr = win_row;
      const bool own_k = (k >= r0 && k < r0 + R);
      for (int peer = 0; peer < CL; ++peer) {
        ClusterTables<T, NBP, MAXCL>* rt = cluster.map_shared_rank(&tab, peer);
        if (tid == 0) { rt->score[par][cta] = ls; rt->row[par][cta] = lr; }
        if (ls >= 0.0 && tid < nb) rt->vals[par][cta][tid] = slab[(lr - (int)r0) * LDS + tid];
        if (own_k && tid >= 32 && tid < 32 + nb) rt->rowk[par][tid - 32] = slab[(k - (int)r0) * LDS + tid - 32];
      }
    }
    cluster.sync();   // release / acquire at cluster scope: the remote stores are visible
    if (tid < 32) {
      double gb = -1.0;
      int grow = INT_MAX, gw = -1;
      for (int w = tid; w < CL; w += 32) {
        const double s2 = tab.score[par][w];
        const int r2 = tab.row[par][w];
        if (s2 >= 0.0 && (s2 > gb || (s2 == gb && r2 < grow))) { gb = s2; grow = r2; gw = w; }
      }
      for (int off = 16; off > 0; off >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, gb, off);
        const int orow = __shfl_xor_sync(0xffffffffu, grow, off);
        const int ow = __shfl_xor_sync(0xffffffffu, gw, off);
        if (ob > gb || (ob == gb && orow < grow)) { gb = ob; grow = orow; gw = ow; }
      }
      if (gw < 0) { gb = 0.0; grow = k; }
      if (tid == 0) { win_score = gb; win_row = grow; }
      if (tid < nb && gw >= 0) { prow[tid] = tab.vals[par][gw][tid]; krow[tid] = tab.rowk[par][tid]; }
    }
    __syncthreads();
    const double gscore = win_score;
    const int piv = win_row;
    if (cta == 0 && tid == 0) {
      ipiv[k] = (int)(row_base + piv + 1);
      if (gscore == 0.0) atomicMin(info, (int)(col_base + k + 1));
    }
    if (gscore != 0.0) {
      if (piv != k) {
        if (k >= r0 && k < r0 + R && tid < nb) slab[(k - (int)r0) * LDS + tid] = prow[tid];
        if (piv >= r0 && piv < r0 + R && tid < nb) slab[(piv - (int)r0) * LDS + tid] = krow[tid];
      }
      __syncthreads();
      const T pv = prow[k];
      for (int r = tid; r < R; r += 256) {
        if (r0 + r <= k) continue;
        T* row = slab + r * LDS;
        const T l = sc_div<T>(row[k], pv);
        row[k] = l;
        for (int j = k + 1; j < nb; ++j) sc_fnma<T>(row[j], l, prow[j]);
      }
    }
    __syncthreads();
  }
  for (int c = 0; c < nb; ++c)
    for (int r = tid; r < R; r += 256) A[(r0 + r) + (int64_t)c * lda] = slab[r * LDS + c];
  cluster.sync();   // nobody exits while a peer may still write into its tables
}

// ---- row interchanges as a permutation -----------------------------------------------------------------------------------
// Simulate ipiv[k0 .. k0+ns) (1-based global rows, relative base row_base = k0) on an index array held in shared memory.
// Output: src_top[k] = source row of destination row k (k < ns), and the list of displaced destinations r >= ns with
// their sources (which all lie in [0, ns)); rows are relative to k0.
__global__ void __launch_bounds__(1024) perm_build_kernel(const int* __restrict__ ipiv, int64_t k0, int ns, int mrows,
                                                          int* __restrict__ src_top, int* __restrict__ disp_dst,
                                                          int* __restrict__ disp_src, int* __restrict__ disp_count) {
  extern __shared__ int idx[];
  __shared__ int count;
  for (int r = threadIdx.x; r < mrows; r += 1024) idx[r] = r;
  if (threadIdx.x == 0) count = 0;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 0; k < ns; ++k) {
      const int p = ipiv[k0 + k] - 1 - (int)k0;
      if (p != k && p >= 0 && p < mrows) { const int t = idx[k]; idx[k] = idx[p]; idx[p] = t; }
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < ns; k += 1024) src_top[k] = idx[k];
  for (int r = ns + threadIdx.x; r < mrows; r += 1024)
    if (idx[r] != r) { const int slot = atomicAdd(&count, 1); disp_dst[slot] = r; disp_src[slot] = idx[r]; }
  __syncthreads();
  if (threadIdx.x == 0) *disp_count = count;
}
// the same simulation in global memory for panels taller than shared memory allows (slow path, one thread)
__global__ void perm_build_global_kernel(const int* __restrict__ ipiv, int64_t k0, int ns, int mrows, int* __restrict__ idx,
                                         int* __restrict__ src_top, int* __restrict__ disp_dst, int* __restrict__ disp_src,
                                         int* __restrict__ disp_count) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  for (int r = 0; r < mrows; ++r) idx[r] = r;
  for (int k = 0; k < ns; ++k) {
    const int p = ipiv[k0 + k] - 1 - (int)k0;
    if (p != k && p >= 0 && p < mrows) { const int t = idx[k]; idx[k] = idx[p]; idx[p] = t; }
  }
  int count = 0;
  for (int k = 0; k < ns; ++k) src_top[k] = idx[k];
  for (int r = ns; r < mrows; ++r)
    if (idx[r] != r) { disp_dst[count] = r; disp_src[count] = idx[r]; ++count; }
  *disp_count = count;
}

// step A: W[k, c] = A[src_top[k], c]          step B: A[dst_i, c] = A[src_i, c]          step C: A[k, c] = W[k, c]
template <typename T>
__global__ void __launch_bounds__(256) perm_gather_kernel(int ns, int64_t ncols, const int* __restrict__ src_top, const T* __restrict__ A,
                                                          int64_t lda, T* __restrict__ W) {
  const int k = blockIdx.x * 256 + threadIdx.x;
  if (k >= ns) return;
  const int s = src_top[k];
  for (int64_t c = blockIdx.y; c < ncols; c += gridDim.y) W[k + c * (int64_t)ns] = A[s + c * lda];
}
template <typename T>
__global__ void __launch_bounds__(256) perm_displace_kernel(const int* __restrict__ count, int64_t ncols, const int* __restrict__ dst,
                                                            const int* __restrict__ src, T* __restrict__ A, int64_t lda) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= *count) return;
  const int d = dst[i], s = src[i];
  for (int64_t c = blockIdx.y; c < ncols; c += gridDim.y) A[d + c * lda] = A[s + c * lda];
}
template <typename T>
__global__ void __launch_bounds__(256) perm_scatter_kernel(int ns, int64_t ncols, const T* __restrict__ W, T* __restrict__ A, int64_t lda) {
  const int k = blockIdx.x * 256 + threadIdx.x;
  if (k >= ns) return;
  for (int64_t c = blockIdx.y; c < ncols; c += gridDim.y) A[k + c * lda] = W[k + c * (int64_t)ns];
}

struct GetrfCtx {
  int sms = 0;
  int* src_top = nullptr; int* disp_dst = nullptr; int* disp_src = nullptr; int* disp_count = nullptr; int* idx_global = nullptr;
  unsigned char* panel_scratch = nullptr;
  void* gslab = nullptr;   // global-memory slab for panels taller than shared memory holds (nullptr: not needed)
  void* W = nullptr; size_t w_bytes = 0;
  size_t max_dyn_smem = 0;
};

template <typename T>
size_t panel_scratch_bytes(int G) {
  constexpr int NBP = Leaf<T>::NB;
  const size_t per = sizeof(PanelCand<T, NBP>) * (size_t)G + sizeof(T) * NBP;
  return 64 + 2 * ((per + 63) / 64 * 64);
}

// apply the interchanges ipiv[k0 .. k0+ns) to columns [c0, c0 + ncols) of A (rows k0 .. m)
template <typename T>
int apply_pivots(const GetrfProblem& p, const GetrfCtx& cx, int64_t k0, int64_t ns, int64_t c0, int64_t ncols, cudaStream_t s) {
  if (ns <= 0 || ncols <= 0) return 0;
  const int64_t mrows = p.m - k0;
  if ((size_t)mrows * sizeof(int) + 2048 <= cx.max_dyn_smem) {
    perm_build_kernel<<<1, 1024, (size_t)mrows * sizeof(int), s>>>(p.dipiv, k0, (int)ns, (int)mrows, cx.src_top, cx.disp_dst, cx.disp_src, cx.disp_count);
  } else {
    perm_build_global_kernel<<<1, 32, 0, s>>>(p.dipiv, k0, (int)ns, (int)mrows, cx.idx_global, cx.src_top, cx.disp_dst, cx.disp_src, cx.disp_count);
  }
  count_launch();
  B200_CUDA_TRY(cudaGetLastError());
  T* A = (T*)p.A + k0;   // rows relative to k0
  const int64_t chunk = std::max<int64_t>(1, (int64_t)(cx.w_bytes / (sizeof(T) * (size_t)ns)));
  const unsigned gx = (unsigned)((ns + 255) / 256);
  for (int64_t c = 0; c < ncols; c += chunk) {
    const int64_t nc = std::min<int64_t>(chunk, ncols - c);
    const unsigned gy = (unsigned)std::min<int64_t>(nc, 4096);
    T* Ac = A + (c0 + c) * p.lda;
    perm_gather_kernel<T><<<dim3(gx, gy), 256, 0, s>>>((int)ns, nc, cx.src_top, Ac, p.lda, (T*)cx.W);
    perm_displace_kernel<T><<<dim3(gx, gy), 256, 0, s>>>(cx.disp_count, nc, cx.disp_dst, cx.disp_src, Ac, p.lda);
    perm_scatter_kernel<T><<<dim3(gx, gy), 256, 0, s>>>((int)ns, nc, (const T*)cx.W, Ac, p.lda);
    count_launch(3);
    B200_CUDA_TRY(cudaGetLastError());
  }
  return 0;
}

template <typename T>
int launch_panel(const GetrfProblem& p, const GetrfCtx& cx, int64_t j0, int64_t nc, cudaStream_t s) {
  constexpr int NBP = Leaf<T>::NB;
  const int64_t mrows = p.m - j0;
  static const bool use_cluster = [] { const char* e = getenv("B200BLAS_GETF2"); return e && e[0] == 'c'; }();   // DRAFT, opt-in
  if (use_cluster) {
    constexpr int MAXCL = 16;
    int cl = (int)std::min<int64_t>(MAXCL, (mrows + 255) / 256);
    if (cl < 1) cl = 1;
    int rpc = (int)((mrows + cl - 1) / cl);
    const size_t smem_c = (size_t)rpc * (NBP + 1) * sizeof(T);
    if (smem_c + sizeof(ClusterTables<T, NBP, MAXCL>) + 4096 <= cx.max_dyn_smem) {
      auto kern = getf2_cluster_kernel<T, NBP, MAXCL>;
      B200_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c));
      B200_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(cl); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem_c; cfg.stream = s;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      B200_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, mrows, (int)nc, rpc, (T*)p.A + j0 + j0 * p.lda, p.lda, p.dipiv + j0, j0, p.dinfo, j0));
      count_launch();
      return 0;
    }   // taller than 16 CTAs' shared memory: the cooperative-grid kernel below
  }
  // about one panel row per thread; the candidate reduce reads one candidate per thread, so G <= 256 CTAs
  int G = (int)std::min<int64_t>(std::min(cx.sms, 256), (mrows + 255) / 256);
  if (G < 1) G = 1;
  int rows_per_cta = (int)((mrows + G - 1) / G);
  size_t smem = (size_t)rows_per_cta * (NBP + 1) * sizeof(T);
  T* gslab = nullptr;
  if (smem + 4096 > cx.max_dyn_smem) {   // taller than G slabs of shared memory: the slab moves to global memory
    if (!cx.gslab) return (int)cudaErrorInvalidConfiguration;
    gslab = (T*)cx.gslab;
    smem = 0;
  }
  B200_SET_MAX_DYN_SMEM_ONCE((getf2_panel_kernel<T, NBP>), cx.max_dyn_smem - 4096);   // the largest slab any panel may need
  B200_CUDA_TRY(cudaMemsetAsync(cx.panel_scratch, 0, 64, s));
  T* A = (T*)p.A + j0 + j0 * p.lda;
  int nb = (int)nc;
  int64_t lda = p.lda, row_base = j0, col_base = j0;
  int* ipiv = p.dipiv + j0;
  int* info = p.dinfo;
  unsigned char* scratch = cx.panel_scratch;
  int64_t mr = mrows;
  void* args[] = {&mr, &nb, &rows_per_cta, &A, &lda, &ipiv, &row_base, &info, &col_base, &scratch, &gslab};
  B200_CUDA_TRY(cudaLaunchCooperativeKernel((const void*)getf2_panel_kernel<T, NBP>, dim3(G), dim3(256), args, smem, s));
  count_launch();
  return 0;
}

// factor columns [j0, j0 + nc) over rows [j0, m); requires j0 + nc <= min(m, n)
template <typename T>
int getrf_rec(const GetrfProblem& p, const GetrfCtx& cx, int64_t j0, int64_t nc, cudaStream_t s) {
  constexpr int NBP = Leaf<T>::NB;
  if (nc <= NBP) return launch_panel<T>(p, cx, j0, nc, s);
  const int64_t n1 = split_point(nc, NBP), n2 = nc - n1;
  T* A = (T*)p.A;
  B200_CUDA_TRY(getrf_rec<T>(p, cx, j0, n1, s));
  B200_CUDA_TRY(apply_pivots<T>(p, cx, j0, n1, j0 + n1, n2, s));          // interchanges of the left half -> right half
  TriProblem t;   // A12 := L11^-1 A12   (PartialPivLU.h:490)
  t.type = p.type; t.left = 1; t.uplo = UPLO_LOWER; t.op = OP_N; t.unit = 1; t.m = n1; t.n = n2;
  t.alpha[0] = 1.0; t.alpha[1] = 0.0;
  t.A = A + j0 + j0 * p.lda; t.lda = p.lda; t.B = A + j0 + (j0 + n1) * p.lda; t.ldb = p.lda;
  B200_CUDA_TRY(launch_trsm(t, s));
  const int64_t mlow = p.m - j0 - n1;
  if (mlow > 0) {   // A22 -= A21 * A12   (PartialPivLU.h:492)
    GemmProblem g;
    g.type = p.type; g.opa = OP_N; g.opb = OP_N; g.m = mlow; g.n = n2; g.k = n1;
    g.alpha[0] = -1.0; g.alpha[1] = 0.0; g.beta[0] = 1.0; g.beta[1] = 0.0;
    g.A = A + (j0 + n1) + j0 * p.lda; g.lda = p.lda;
    g.B = t.B; g.ldb = p.lda;
    g.C = A + (j0 + n1) + (j0 + n1) * p.lda; g.ldc = p.lda;
    B200_CUDA_TRY(run_gemm_device(g, s, B200BLAS_AUTO));
  }
  B200_CUDA_TRY(getrf_rec<T>(p, cx, j0 + n1, n2, s));
  return apply_pivots<T>(p, cx, j0 + n1, n2, j0, n1, s);                 // interchanges of the right half -> left half
}

template <typename T>
int getrf_typed(const GetrfProblem& p, cudaStream_t s) {
  GetrfCtx cx;
  int dev = 0;
  B200_CUDA_TRY(cudaGetDevice(&dev));
  B200_CUDA_TRY(cudaDeviceGetAttribute(&cx.sms, cudaDevAttrMultiProcessorCount, dev));
  int optin = 0;
  B200_CUDA_TRY(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  cx.max_dyn_smem = (size_t)optin;
  B200_SET_MAX_DYN_SMEM_ONCE(perm_build_kernel, optin - 1024);
  const int64_t size = std::min(p.m, p.n);
  // workspace: permutation lists (4 * size ints + m ints for the slow path), panel scratch, gather buffer
  const size_t ints = (size_t)size * 3 + 16 + (size_t)p.m;
  const size_t ps = panel_scratch_bytes<T>(cx.sms);
  cx.w_bytes = std::min<size_t>((size_t)64 << 20, std::max<size_t>((size_t)1 << 20, (size_t)size * (size_t)p.n * sizeof(T)));
  // a gather chunk must hold at least one column of `size / 2` rows
  cx.w_bytes = std::max(cx.w_bytes, (size_t)size * sizeof(T));
  // tall panels (the first panel is the tallest): G slabs of shared memory hold G * (optin - 4096) bytes
  constexpr int NBP = Leaf<T>::NB;
  const int gmax = std::min(cx.sms, 256);
  const int64_t rpc0 = (p.m + gmax - 1) / gmax;
  size_t gslab_bytes = 0;
  if ((size_t)rpc0 * (NBP + 1) * sizeof(T) + 4096 > cx.max_dyn_smem) gslab_bytes = (size_t)gmax * (size_t)rpc0 * (NBP + 1) * sizeof(T);
  const size_t total = (ints * sizeof(int) + 255) / 256 * 256 + (ps + 255) / 256 * 256 + (cx.w_bytes + 255) / 256 * 256 + gslab_bytes;
  unsigned char* ws = nullptr;
  B200_CUDA_TRY(cudaMallocAsync((void**)&ws, total, s));
  int* ip = (int*)ws;
  cx.src_top = ip; cx.disp_dst = ip + size; cx.disp_src = ip + 2 * size; cx.disp_count = ip + 3 * size; cx.idx_global = ip + 3 * size + 16;
  cx.panel_scratch = ws + (ints * sizeof(int) + 255) / 256 * 256;
  cx.W = cx.panel_scratch + (ps + 255) / 256 * 256;
  if (gslab_bytes) cx.gslab = (unsigned char*)cx.W + (cx.w_bytes + 255) / 256 * 256;
  int e = getrf_rec<T>(p, cx, 0, size, s);
  if (!e && p.n > size) {   // wide matrix: the columns right of the square part (LAPACK semantics)
    e = apply_pivots<T>(p, cx, 0, size, size, p.n - size, s);
    if (!e) {
      TriProblem t;
      t.type = p.type; t.left = 1; t.uplo = UPLO_LOWER; t.op = OP_N; t.unit = 1; t.m = size; t.n = p.n - size;
      t.alpha[0] = 1.0; t.alpha[1] = 0.0;
      t.A = p.A; t.lda = p.lda; t.B = (T*)p.A + size * p.lda; t.ldb = p.lda;
      e = launch_trsm(t, s);
    }
  }
  cudaFreeAsync(ws, s);
  return e;
}

}  // namespace

// *p.dinfo must hold INT_MAX on entry; on exit it is the smallest failing 1-based index, or still INT_MAX
int launch_potrf(const PotrfProblem& p, cudaStream_t s) {
  if (p.n <= 0) return 0;
  note_variant("potrf_recursive_leaf+trsm+syrk");
  switch (p.type) {
    case TY_S: return potrf_rec<float>(p, 0, p.n, s);
    case TY_D: return potrf_rec<double>(p, 0, p.n, s);
    case TY_C: return potrf_rec<float2>(p, 0, p.n, s);
    default: return potrf_rec<double2>(p, 0, p.n, s);
  }
}

int launch_getrf(const GetrfProblem& p, cudaStream_t s) {
  if (p.m <= 0 || p.n <= 0) return 0;
  note_variant("getrf_recursive_panel+trsm+gemm");
  switch (p.type) {
    case TY_S: return getrf_typed<float>(p, s);
    case TY_D: return getrf_typed<double>(p, s);
    case TY_C: return getrf_typed<float2>(p, s);
    default: return getrf_typed<double2>(p, s);
  }
}

}  // namespace b200
