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
#include <type_traits>

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

// ---- forced variant of the recursive path (B200BLAS_POTRF=rec B200BLAS_POTF2=cta; 115 tests green on B200, pass 1b) -----------
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

// ---- Cholesky of one diagonal block of order <= NBL (128; 64 for complex double) by ONE CTA -----------------------------------
// The block (lower-canonical: element (i, j) = A(i,j) for uplo = L, conj(A(j,i)) for uplo = U) lives in shared memory.
// It is factored in panels of 32 columns: thread t owns row c0 + t of the panel in 32 registers; a column costs ONE
// __syncthreads -- the raw (unscaled) entries of the diagonal rows are published, every thread derives the pivot root
// and the scaled multipliers from them itself -- and the part of the block right of the panel is then updated by all
// 256 threads with 4 x 4 register tiles (rows / columns interleaved so that every shared-memory access is conflict
// free).  Rows / columns past d are the identity, so the unrolled loops are valid for every d <= NBL.
template <typename T> struct Blk { static constexpr int NBL = 128; };
template <> struct Blk<double2> { static constexpr int NBL = 64; };

template <typename T, int NBL>
__global__ void __launch_bounds__(256) potf2_block_kernel(int upper, int d, int64_t d0, T* __restrict__ A, int64_t lda, int* __restrict__ info) {
  using R = typename Sc<T>::real;
  constexpr int LDS = NBL + 1;
  extern __shared__ __align__(16) unsigned char blk_smem[];
  T* S = reinterpret_cast<T*>(blk_smem);   // S[i * LDS + j], i >= j
  __shared__ T colbuf[3][64];
  __shared__ R roots[2][32];
  const int tid = threadIdx.x;
  {
    // all loads of a batch are issued before any is consumed (a load-per-iteration loop costs one DRAM latency per element:
    // 64 x 700 cycles for this block)
    constexpr int BATCH = 16;
    for (int base = 0; base < NBL * NBL; base += 256 * BATCH) {
      T vals[BATCH];
#pragma unroll
      for (int q = 0; q < BATCH; ++q) {
        const int idx = base + q * 256 + tid;
        int i, j;
        if (upper) { j = idx % NBL; i = idx / NBL; } else { i = idx % NBL; j = idx / NBL; }   // coalesced either way
        vals[q] = (i == j) ? sc_one<T>() : Sc<T>::zero();
        if (idx < NBL * NBL && i < d && j <= i) vals[q] = upper ? Sc<T>::conj(A[j + (int64_t)i * lda]) : A[i + (int64_t)j * lda];
      }
#pragma unroll
      for (int q = 0; q < BATCH; ++q) {
        const int idx = base + q * 256 + tid;
        int i, j;
        if (upper) { j = idx % NBL; i = idx / NBL; } else { i = idx % NBL; j = idx / NBL; }
        if (idx < NBL * NBL) S[i * LDS + j] = vals[q];
      }
    }
  }
  __syncthreads();
  bool ok = true;   // uniform over the CTA
  // colbuf[p][j]: raw (unscaled) entries of the current column in the diagonal rows, three buffers in rotation; [32, 64) stays
  // zero so that the rotated update below may read L(c0 + k + j, .) for every j without a bounds test
  for (int idx = tid; idx < 3 * 64; idx += 256) colbuf[idx / 64][idx % 64] = Sc<T>::zero();
  __syncthreads();
  for (int c0 = 0; c0 < NBL && c0 < d && ok; c0 += 32) {
    const int t = tid;
    const bool rowact = t < NBL - c0;
    // the thread's row of the panel, ROTATED: a[j] = element (c0 + t, c0 + k + j) at column step k.  The column loop is a
    // runtime loop over a compact body (the unrolled version was 120 KB of straight-line code and instruction-fetch bound)
    T a[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) a[j] = rowact ? S[(c0 + t) * LDS + c0 + j] : Sc<T>::zero();
    if (t < 32) colbuf[0][t] = a[0];
    __syncthreads();
    // Critical path of a column = barrier -> pivot x -> 1/x -> update of the NEXT column's entry -> publish -> barrier:
    // the update uses w = a(i,k) / x against the RAW column (a(i,j) -= w * conj(a(j,k)), the same product as
    // L(i,k) conj(L(j,k))), so no square root is on that path -- the column is stored unscaled, the pivots are kept, and
    // the 32 roots are taken in parallel after the loop (one thread each), followed by one scaling pass over the panel.
    R x = sc_real<T>(colbuf[0][0]);
    R inv_x = fast_rcp(x);
    int ncols = 32;   // columns of this panel that were factored (all of them unless a pivot failed)
#pragma unroll 1
    for (int k = 0; k < 32; ++k) {
      if (x <= (R)0) {   // a NaN pivot continues, as in the reference (LLT.h:316-317)
        if (tid == 0) atomicMin(info, (int)(d0 + c0 + k + 1));
        ok = false;
        ncols = k;
        break;
      }
      const T* col = colbuf[k % 3];
      const bool below = rowact && t > k;
      const T a0 = a[0];
      const T w = below ? sc_scale<T>(a0, inv_x) : Sc<T>::zero();
      T v1 = a[1];
      sc_fnma<T>(v1, w, Sc<T>::conj(col[k + 1]));
      R xn = (R)1, inv_xn = (R)1;
      if (k + 1 < 32) {
        if (t < 32 && t > k) colbuf[(k + 1) % 3][t] = v1;
        __syncthreads();
        xn = sc_real<T>(colbuf[(k + 1) % 3][k + 1]);
        inv_xn = fast_rcp(xn);
      }
      if (rowact && t >= k) S[(c0 + t) * LDS + c0 + k] = (t == k) ? sc_from_real<T>(x) : a0;   // unscaled; the diagonal keeps the pivot
      a[0] = v1;
#pragma unroll
      for (int j = 2; j < 32; ++j) {
        T v = a[j];
        sc_fnma<T>(v, w, Sc<T>::conj(col[k + j]));   // raw a(c0 + k + j, c0 + k); zero past the panel
        a[j - 1] = v;
      }
      a[31] = Sc<T>::zero();
      x = xn; inv_x = inv_xn;
    }
    __syncthreads();
    // roots of the panel's pivots (one thread each), then L(i, k) = a(i, k) / sqrt(x_k) for the whole panel
    if (tid < ncols) {
      R l, rl;
      pivot_roots(sc_real<T>(S[(c0 + tid) * LDS + c0 + tid]), l, rl);
      roots[0][tid] = l; roots[1][tid] = rl;
    }
    __syncthreads();
    if (rowact) {
      T* srow = S + (c0 + t) * LDS + c0;
      for (int k = 0; k < ncols && k <= t; ++k) srow[k] = (k == t) ? sc_from_real<T>(roots[0][k]) : sc_scale<T>(srow[k], roots[1][k]);
    }
    __syncthreads();
    const int c1 = c0 + 32;
    if (!ok || c1 >= NBL || c1 >= d) continue;
    // S[i][j] -= sum_k L[i][c0 + k] conj(L[j][c0 + k]) for i >= j >= c1 (identity rows past d have zero L entries)
    const int nt = (NBL - c1) / 4;
    for (int tile = tid; tile < nt * nt; tile += 256) {
      const int ti = tile / nt, tj = tile % nt;
      T acc[4][4];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int i = c1 + ti + r * nt, j = c1 + tj + c * nt;
          acc[r][c] = (i >= j) ? S[i * LDS + j] : Sc<T>::zero();
        }
      const T* li = S + (c1 + ti) * LDS + c0;
      const T* lj = S + (c1 + tj) * LDS + c0;
#pragma unroll 4
      for (int k = 0; k < 32; ++k) {
        T vi[4], vj[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) { vi[r] = li[r * nt * LDS + k]; vj[r] = Sc<T>::conj(lj[r * nt * LDS + k]); }
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) sc_fnma<T>(acc[r][c], vi[r], vj[c]);
      }
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int i = c1 + ti + r * nt, j = c1 + tj + c * nt;
          if (i >= j) S[i * LDS + j] = acc[r][c];
        }
    }
    __syncthreads();
  }
  for (int idx = tid; idx < NBL * NBL; idx += 256) {
    int i, j;
    if (upper) { j = idx % NBL; i = idx / NBL; } else { i = idx % NBL; j = idx / NBL; }
    if (i < d && j <= i) {
      const T v = S[i * LDS + j];
      if (upper) A[j + (int64_t)i * lda] = Sc<T>::conj(v); else A[i + (int64_t)j * lda] = v;
    }
  }
}

// ---- auxiliary stream + events of the look-ahead factorizations (one set per host thread and device) -------------------------
struct LookAhead {
  cudaStream_t sp = nullptr;   // panel chain, highest priority: its small kernels slip in between the CTAs of the big update
  cudaEvent_t e_in = nullptr, e_panel = nullptr, e_rest = nullptr, e_out = nullptr;
  int dev = -1;
  int init() {
    int cur = 0;
    B200_CUDA_TRY(cudaGetDevice(&cur));
    if (sp && dev == cur) return 0;
    release();
    int lo = 0, hi = 0;
    B200_CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    B200_CUDA_TRY(cudaStreamCreateWithPriority(&sp, cudaStreamNonBlocking, hi));
    for (cudaEvent_t* e : {&e_in, &e_panel, &e_rest, &e_out}) B200_CUDA_TRY(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    dev = cur;
    return 0;
  }
  void release() {
    if (sp) cudaStreamDestroy(sp);
    for (cudaEvent_t* e : {&e_in, &e_panel, &e_rest, &e_out}) { if (*e) cudaEventDestroy(*e); *e = nullptr; }
    sp = nullptr; dev = -1;
  }
};
thread_local LookAhead t_look;

bool lookahead_enabled() {
  static const bool v = [] { const char* e = getenv("B200BLAS_LOOKAHEAD"); return !(e && e[0] == '0'); }();
  return v;
}

// C[r0.., c0..] -= L[r0.., j..j+jb) * L[c0.., j..j+jb)^H in lower-canonical coordinates (r0 == c0: only the referenced
// triangle of the window is touched); for uplo = U the same update on the stored transpose.
template <typename T>
int potrf_update(const PotrfProblem& p, int64_t j, int64_t jb, int64_t r0, int64_t nr, int64_t c0, int64_t nc, cudaStream_t s) {
  if (nr <= 0 || nc <= 0) return 0;
  const bool cplx = sizeof(T) != sizeof(typename Sc<T>::real);
  T* A = (T*)p.A;
  GemmProblem g;
  g.type = p.type; g.alpha[0] = -1.0; g.alpha[1] = 0.0; g.beta[0] = 1.0; g.beta[1] = 0.0;
  g.k = jb; g.uplo = p.uplo; g.herm = cplx ? 1 : 0; g.lda = g.ldb = g.ldc = p.lda;
  if (p.uplo != UPLO_UPPER) {
    g.opa = OP_N; g.opb = OP_C; g.m = nr; g.n = nc;
    g.A = A + r0 + j * p.lda; g.B = A + c0 + j * p.lda; g.C = A + r0 + c0 * p.lda;
  } else {
    g.opa = OP_C; g.opb = OP_N; g.m = nc; g.n = nr;
    g.A = A + j + c0 * p.lda; g.B = A + j + r0 * p.lda; g.C = A + c0 + r0 * p.lda;
  }
  return run_gemm_device(g, s, B200BLAS_AUTO);
}

// Right-looking blocked Cholesky with one step of look-ahead (the blocked loop of LLT.h:330-360 with blockSize = NBL):
//   panel chain (stream sp):  factor A11 (one CTA) -> A21 := A21 L11^-H (one substitution-leaf launch) -> update of the NEXT
//                             block column only -> factor the next A11 ...
//   bulk (the caller's stream): A22 -= A21 A21^H on everything right of the next block column ("bottleneck", LLT.h:357),
// so that the latency-bound panel chain of step j+1 runs underneath the tensor-pipe update of step j.
template <typename T>
int potrf_blocked(const PotrfProblem& p, cudaStream_t s) {
  constexpr int NBL = Blk<T>::NBL;
  const bool upper = p.uplo == UPLO_UPPER;
  T* A = (T*)p.A;
  const int64_t n = p.n;
  constexpr size_t smem = (size_t)NBL * (NBL + 1) * sizeof(T);
  B200_SET_MAX_DYN_SMEM_ONCE((potf2_block_kernel<T, NBL>), smem);
  // Two-level blocking: the one-CTA block kernel factors NBL columns at a time, but the bulk update uses the last NBO columns as
  // its k dimension: 2 * NBL halves the number of passes over the trailing matrix (a rank-256 update runs at 0.87 of the
  // DMMA pipe, a rank-128 one at ~0.75).  Measured on B200 (profiles/bench_r02/potrf_two_level_pass9.txt), NBO = 128 / 256 /
  // 512: dpotrf 8192 14.84 / 15.06 / 14.08, dpotrf 16384 23.33 / 25.20 / 25.64, spotrf 8192 16.41 / 17.23 / 17.52 TFLOP/s.
  // B200BLAS_POTRF_NB=<multiple of NBL> overrides.
  static const int nbo_env = [] { const char* e = getenv("B200BLAS_POTRF_NB"); return e ? atoi(e) : 0; }();
  const int64_t nbo_want = nbo_env > 0 ? nbo_env : 2 * NBL;
  const int64_t NBO = (nbo_want >= NBL && nbo_want % NBL == 0 && n >= 8 * nbo_want) ? nbo_want : NBL;
  const bool look = lookahead_enabled() && n > 4 * NBO;
  cudaStream_t sp = s;
  LookAhead& la = t_look;
  if (look) {
    B200_CUDA_TRY(la.init());
    sp = la.sp;
    B200_CUDA_TRY(cudaEventRecord(la.e_in, s));
    B200_CUDA_TRY(cudaStreamWaitEvent(sp, la.e_in, 0));
  }
  bool rest_pending = false;
  for (int64_t j = 0; j < n; j += NBO) {
    const int64_t jb = std::min<int64_t>(NBO, n - j), n2 = n - j - jb;
    // panel [j, j + jb): block by block -- factor the diagonal block, solve the rows below it, update the rest of the panel
    for (int64_t jj = j; jj < j + jb; jj += NBL) {
      const int64_t jbb = std::min<int64_t>(NBL, j + jb - jj), below = n - jj - jbb;
      potf2_block_kernel<T, NBL><<<1, 256, smem, sp>>>(upper ? 1 : 0, (int)jbb, jj, A + jj + jj * p.lda, p.lda, p.dinfo);
      count_launch();
      B200_CUDA_TRY(cudaGetLastError());
      if (below <= 0) break;
      TriProblem t;   // A21 := A21 * L11^-H  /  A12 := U11^-H * A12   (LLT.h:356)
      t.type = p.type; t.uplo = p.uplo; t.op = OP_C; t.unit = 0; t.alpha[0] = 1.0; t.alpha[1] = 0.0;
      t.A = A + jj + jj * p.lda; t.lda = p.lda; t.ldb = p.lda;
      if (!upper) { t.left = 0; t.m = below; t.n = jbb; t.B = A + (jj + jbb) + jj * p.lda; }
      else { t.left = 1; t.m = jbb; t.n = below; t.B = A + jj + (jj + jbb) * p.lda; }
      B200_CUDA_TRY(launch_trsm(t, sp));
      const int64_t rp = jj + jbb, ncp = j + jb - rp;   // rest of the panel: columns [rp, j + jb), rows rp..
      if (ncp > 0) B200_CUDA_TRY(potrf_update<T>(p, jj, jbb, rp, n - rp, rp, ncp, sp));
    }
    if (n2 <= 0) break;
    const int64_t r1 = j + jb, nb2 = std::min<int64_t>(NBO, n2), r2 = r1 + nb2, n3 = n - r2;
    if (look) {
      B200_CUDA_TRY(cudaEventRecord(la.e_panel, sp));
      if (rest_pending) B200_CUDA_TRY(cudaStreamWaitEvent(sp, la.e_rest, 0));   // the previous bulk update also touched the next block column
    }
    // next block column: rows r1.., columns [r1, r2)
    B200_CUDA_TRY(potrf_update<T>(p, j, jb, r1, n2, r1, nb2, sp));
    if (n3 > 0) {
      if (look) B200_CUDA_TRY(cudaStreamWaitEvent(s, la.e_panel, 0));
      B200_CUDA_TRY(potrf_update<T>(p, j, jb, r2, n3, r2, n3, s));
      if (look) { B200_CUDA_TRY(cudaEventRecord(la.e_rest, s)); rest_pending = true; }
    }
  }
  if (look) {
    B200_CUDA_TRY(cudaEventRecord(la.e_out, sp));
    B200_CUDA_TRY(cudaStreamWaitEvent(s, la.e_out, 0));
  }
  return 0;
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
  static const bool cta_leaf = [] { const char* e = getenv("B200BLAS_POTF2"); return e && e[0] == 'c'; }();   // forced variant, opt-in
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

// ---- forced variant (B200BLAS_GETF2=cluster; 115 tests green on B200, pass 1b; no faster than the cooperative grid: the slab
// arithmetic, not the barrier, was the bound -- which is what led to the register-resident panel below) ---------------------------
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

// ---- LU panel, register resident, on one thread-block cluster (the default leaf for panels of up to 16384 rows) ----------------
// Every thread owns RPT whole rows of the (<= NBP wide) panel in registers, so the rank-1 update of a column is pure
// register arithmetic against the pivot row (broadcast from shared memory) -- the shared-memory slab kernels above spend
// their time in dependent LDS -> DFMA -> STS chains (4.1 us per column measured in round 1 and again under the cluster
// draft).  Per column: local arg-max (registers, shuffles) -> the CTA's candidate row and, from its owner, row k are
// pushed into the tables of EVERY CTA of the cluster through distributed shared memory -> one hardware cluster barrier
// -> identical reduction everywhere -> the owners of rows k / pivot overwrite their registers with each other's row ->
// scale and update.  Tables are double-buffered by column parity: one cluster barrier per column.
// compile-time loop: the column index of the register panels must be a constant in every iteration, also where the
// body is too large for `#pragma unroll` to be honoured (the row arrays would otherwise move to local memory)
template <int K, int N, typename F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (K < N) {
    f(std::integral_constant<int, K>{});
    static_for<K + 1, N>(f);
  }
}

template <typename T, int NBP>
struct RegPanelTables {
  unsigned long long score[2][16];   // score_key of every CTA's candidate
  int row[2][16];
  T vals[2][16][NBP];
  T rowk[2][NBP];
};

// The column loop is a RUNTIME loop over a compact body: the rows are kept "rotated" -- the active column is always
// element 0 of a thread's row array, and the rank-1 update writes a[j-1] = a[j] - l * u[j], shifting the row left by one --
// so no register index depends on k.  (A fully unrolled version of this kernel is 320 KB of straight-line code; ncu showed 44 %
// of its stall samples as "no instruction": profiles/ncu_r02_lapack_leaves.md.)  What falls off the left end is final: the
// multipliers l of column k go to a shared-memory column buffer (Lbuf), row k of U -- known to every thread: it is the pivot
// row everybody just received -- to a buffer in CTA 0; both are flushed to global memory once, after the last column, so
// that the per-column cluster barrier never has to wait for global stores (the version that stored them at once spent 15 %
// of its stall samples in the barrier's fence).  The interchange of the already-final L part of rows k and pivot (columns
// < k) is done on the Lbuf columns, through distributed shared memory, by one warp.  All reductions are shuffle trees
// (max of the scores, then the smallest row among the lanes that hold the maximum = maxCoeff's "first").
// A pivot candidate is (key, row): key = bit pattern of the score |a| plus one (doubles >= 0 order like their bit patterns;
// 0 = no candidate: no row, or a NaN score, which never compares greater in maxCoeff).  Warp-wide arg-max = three REDUX
// instructions (max of the high words, max of the low words among the lanes that hold it, min of the rows among the lanes
// that hold both) instead of a five-level shuffle tree on 64-bit values.
__device__ __forceinline__ unsigned long long score_key(double sc) {
  return sc >= 0.0 ? (unsigned long long)__double_as_longlong(sc) + 1ull : 0ull;   // false for NaN
}
__device__ __forceinline__ void argmax_tree(unsigned long long& key, int& row) {
  const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
  const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
  const unsigned mlo = __reduce_max_sync(0xffffffffu, hi == mhi ? lo : 0u);
  const unsigned long long m = ((unsigned long long)mhi << 32) | mlo;
  row = (int)__reduce_min_sync(0xffffffffu, (unsigned)((key == m) ? row : INT_MAX));
  key = m;
}

template <typename T, int NBP, int RPT, int THREADS>
__global__ void __launch_bounds__(THREADS)
getf2_reg_kernel(int64_t mrows, int nb, T* __restrict__ A, int64_t lda, int* __restrict__ ipiv, int64_t row_base,
                 int* __restrict__ info, int64_t col_base) {
  constexpr int NW = THREADS / 32, RT = RPT * THREADS;
  static_assert(NW <= 32 && NBP <= 32, "one lane per warp result / per panel column");
  extern __shared__ __align__(16) unsigned char reg_smem[];
  T* Lbuf = reinterpret_cast<T*>(reg_smem);   // Lbuf[c * RT + local row]: multipliers of column c, this CTA's rows
  __shared__ RegPanelTables<T, NBP> tab;
  __shared__ T Ubuf[NBP][NBP];                // CTA 0: Ubuf[k][j] = U(k, k + j)
  __shared__ int ipiv_s[NBP];                 // CTA 0
  __shared__ unsigned long long wbest[NW];
  __shared__ int wrow[NW];
  __shared__ T myrow[NBP], mykrow[NBP];
  cg::cluster_group cluster = cg::this_cluster();
  const int CL = (int)cluster.num_blocks(), cta = (int)cluster.block_rank(), tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  int rq[RPT];
  T a[RPT][NBP];   // a[q][j] = current value of element (row rq[q], column k + j)
#pragma unroll
  for (int q = 0; q < RPT; ++q) {
    rq[q] = cta * RT + q * THREADS + tid;
#pragma unroll
    for (int c = 0; c < NBP; ++c) a[q][c] = (rq[q] < mrows && c < nb) ? A[rq[q] + (int64_t)c * lda] : Sc<T>::zero();
  }
  cluster.sync();   // every CTA's shared memory is live before anybody writes into it remotely
  const int steps = (int)min((int64_t)nb, mrows);
#pragma unroll 1
  for (int k = 0; k < steps; ++k) {
    const int par = k & 1;
    // 1. candidate of this thread -> warp: largest |a(r,k)| among rows >= k, smallest row on ties (maxCoeff keeps the first)
    unsigned long long best = 0ull;
    int brow = INT_MAX;
#pragma unroll
    for (int q = 0; q < RPT; ++q) {
      if (rq[q] >= k && rq[q] < mrows) {
        const unsigned long long sk = score_key(sc_score<T>(a[q][0]));
        if (sk > best) { best = sk; brow = rq[q]; }   // rq[0] < rq[1]: the first row wins a tie
      }
    }
    argmax_tree(best, brow);
    if (lane == 0) { wbest[warp] = best; wrow[warp] = brow; }
    __syncthreads();
    // 2. warp results -> the CTA's candidate (every warp runs the same tree); its owner and the owner of row k publish their rows
    unsigned long long cb = lane < NW ? wbest[lane] : 0ull;
    int crow = lane < NW ? wrow[lane] : INT_MAX;
    argmax_tree(cb, crow);
    const int krow_cta = k / RT;   // the CTA that owns row k
#pragma unroll
    for (int q = 0; q < RPT; ++q) {
      if (cb != 0ull && rq[q] == crow) {
#pragma unroll
        for (int c = 0; c < NBP; ++c) myrow[c] = a[q][c];
      }
      if (rq[q] == k) {
#pragma unroll
        for (int c = 0; c < NBP; ++c) mykrow[c] = a[q][c];
      }
    }
    __syncthreads();
    for (int idx = tid; idx < CL * NBP; idx += THREADS) {
      const int peer = idx / NBP, c = idx % NBP;
      RegPanelTables<T, NBP>* rt = cluster.map_shared_rank(&tab, peer);
      if (cb != 0ull) rt->vals[par][cta][c] = myrow[c];
      if (cta == krow_cta) rt->rowk[par][c] = mykrow[c];
    }
    if (tid < CL) {
      RegPanelTables<T, NBP>* rt = cluster.map_shared_rank(&tab, tid);
      rt->score[par][cta] = cb;
      rt->row[par][cta] = crow;
    }
    cluster.sync();   // release / acquire at cluster scope: the remote stores of this column are visible
    // 3. identical reduction of the CL candidates in every warp
    unsigned long long gb = 0ull;
    int grow = INT_MAX;
    if (lane < CL) {
      gb = tab.score[par][lane];
      if (gb != 0ull) grow = tab.row[par][lane];
    }
    const int my_row = grow;
    argmax_tree(gb, grow);
    const bool nonzero = gb > 1ull;       // key 1 = score 0: zero (or, key 0, all-NaN) pivot column: recorded and skipped (PartialPivLU.h:396-401)
    const int gw = __ffs(__ballot_sync(0xffffffffu, my_row == grow && grow != INT_MAX)) - 1;   // the CTA that holds the winner
    const int piv = (nonzero && gw >= 0) ? grow : k;
    const bool have = nonzero && gw >= 0;
    const T* urow = have ? tab.vals[par][gw] : tab.rowk[par];   // row k of U, in rotated coordinates (urow[j] = U(k, k + j))
    if (cta == 0 && tid == 0) {
      ipiv_s[k] = (int)(row_base + piv + 1);
      if (!have) atomicMin(info, (int)(col_base + k + 1));
    }
    // row k is final: U(k, k..) is kept in CTA 0; the finished L part (columns < k) of rows k and piv is interchanged in the
    // Lbuf columns of their owners, by one warp (lane c = column c); its loads are issued here and consumed after the update
    const bool helper = cta == 0 && warp == 1;
    const bool swap_l = helper && piv != k && lane < k;
    T t1 = Sc<T>::zero(), t2 = Sc<T>::zero();
    T* lk = nullptr;
    T* lp = nullptr;
    if (helper && lane < NBP) Ubuf[k][lane] = urow[lane];
    if (swap_l) {
      lk = cluster.map_shared_rank(Lbuf, k / RT) + (size_t)lane * RT + k % RT;
      lp = cluster.map_shared_rank(Lbuf, piv / RT) + (size_t)lane * RT + piv % RT;
      t1 = *lk; t2 = *lp;
    }
    const T inv_pv = have ? sc_fast_recip<T>(urow[0]) : Sc<T>::zero();
#pragma unroll
    for (int q = 0; q < RPT; ++q) {
      // 4. interchange rows k and piv (PartialPivLU.h:384-388): the owner of position piv takes over the old row k
      if (piv != k && rq[q] == piv) {
#pragma unroll
        for (int c = 0; c < NBP; ++c) a[q][c] = tab.rowk[par][c];
      }
      // 5. scale the column below the pivot, keep it, update the rest of the row and rotate it (PartialPivLU.h:392, :404-405).
      //    Rows that take no part (<= k, or past the panel) get l = 0; their registers are dead from here on.
      const bool act = rq[q] > k && rq[q] < mrows;
      T l = Sc<T>::zero();
      if (act) {
        l = have ? Sc<T>::mul(a[q][0], inv_pv) : a[q][0];
        Lbuf[(size_t)k * RT + q * THREADS + tid] = l;
        if (!have) l = Sc<T>::zero();
      }
#pragma unroll
      for (int j = 1; j < NBP; ++j) {
        T v = a[q][j];
        sc_fnma<T>(v, l, urow[j]);
        a[q][j - 1] = v;
      }
      a[q][NBP - 1] = Sc<T>::zero();
    }
    if (swap_l) { *lk = t2; *lp = t1; }
  }
  cluster.sync();   // every multiplier and every interchange has landed in its Lbuf
  // flush: L part of this CTA's rows (coalesced), and from CTA 0 the U rows and the pivots
#pragma unroll
  for (int q = 0; q < RPT; ++q) {
    if (rq[q] < mrows) {
      const int ncol = min(rq[q], steps);   // columns c < r (and < steps) hold L(r, c)
      for (int c = 0; c < ncol; ++c) A[rq[q] + (int64_t)c * lda] = Lbuf[(size_t)c * RT + q * THREADS + tid];
    }
  }
  if (cta == 0) {
    for (int idx = tid; idx < steps * NBP; idx += THREADS) {
      const int k = idx / NBP, j = idx % NBP;
      if (k + j < nb) A[k + (int64_t)(k + j) * lda] = Ubuf[k][j];
    }
    if (tid < steps) ipiv[tid] = ipiv_s[tid];
  }
  cluster.sync();   // nobody exits while a peer may still access its shared memory
}

// ---- two chained 32-column phases in ONE launch (panel width NBP < nb <= 2 NBP, at least 2 NBP rows) ----------------------------
// The lowest inner node of the panel recursion -- interchanges of the left leaf applied to the right one, a 32 x 32 row
// solve, a rank-32 update, the right leaf, and the right leaf's interchanges applied back to the left one: five launches of
// 7 - 40 us each on the critical path of ?getrf_ -- is folded into the leaf kernel: after phase 0 every thread GATHERS its
// row of the next 32 columns from the position the interchanges so far assign to it (nothing has been written there yet),
// every CTA forms U12 = L11^-1 A12 in shared memory (L11 read from CTA 0's column buffer through DSMEM, rotated forward
// substitution by one warp), every thread subtracts L(r, 0:32) U12 from its row using ITS OWN multipliers, still in the column
// buffer, phase 0 is flushed, and phase 1 runs the same column loop.  At the end CTA 0 applies phase 1's interchanges to the
// L part of phase 0's columns (the same reverse-order bookkeeping as perm_apply_direct_kernel).
template <typename T, int NBP, int RPT, int THREADS>
__global__ void __launch_bounds__(THREADS)
getf2_reg2_kernel(int64_t mrows, int nb, T* __restrict__ A, int64_t lda, int* __restrict__ ipiv, int64_t row_base,
                  int* __restrict__ info, int64_t col_base) {
  constexpr int NW = THREADS / 32, RT = RPT * THREADS;
  static_assert(NW <= 32 && NBP <= 32 && THREADS >= 2 * NBP, "one lane per warp result / per panel column");
  extern __shared__ __align__(16) unsigned char reg_smem[];
  T* Lbuf = reinterpret_cast<T*>(reg_smem);   // Lbuf[c * RT + local row]: multipliers of column c of the CURRENT phase
  __shared__ RegPanelTables<T, NBP> tab;
  __shared__ T Ubuf[2 * NBP][NBP];            // CTA 0: Ubuf[K][j] = U(K, K + j) within K's phase
  __shared__ T U12[NBP][NBP];                 // every CTA: rows 0 .. NBP-1 of the columns of phase 1
  __shared__ T L11s[NBP][NBP + 1];            // every CTA: unit-lower block of phase 0
  __shared__ int ipiv_s[2 * NBP];             // CTA 0: 1-based global rows
  __shared__ int pivrel[2 * NBP];             // every CTA: pivot rows relative to the panel
  __shared__ unsigned long long wbest[NW];
  __shared__ int wrow[NW];
  __shared__ T myrow[NBP], mykrow[NBP];
  cg::cluster_group cluster = cg::this_cluster();
  const int CL = (int)cluster.num_blocks(), cta = (int)cluster.block_rank(), tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int nb1 = nb - NBP;   // columns of phase 1
  int rq[RPT];
  T a[RPT][NBP];   // a[q][j] = current value of element (row rq[q], column K + j)
#pragma unroll
  for (int q = 0; q < RPT; ++q) {
    rq[q] = cta * RT + q * THREADS + tid;
#pragma unroll
    for (int c = 0; c < NBP; ++c) a[q][c] = rq[q] < mrows ? A[rq[q] + (int64_t)c * lda] : Sc<T>::zero();
  }
  cluster.sync();   // every CTA's shared memory is live before anybody writes into it remotely

  // one phase: columns K0 .. K0 + nsteps - 1 of the panel (the column loop of getf2_reg_kernel with a row / column offset)
  auto run_phase = [&](const int K0, const int nsteps) {
#pragma unroll 1
    for (int kk = 0; kk < nsteps; ++kk) {
      const int K = K0 + kk, par = kk & 1;
      unsigned long long best = 0ull;
      int brow = INT_MAX;
#pragma unroll
      for (int q = 0; q < RPT; ++q) {
        if (rq[q] >= K && rq[q] < mrows) {
          const unsigned long long sk = score_key(sc_score<T>(a[q][0]));
          if (sk > best) { best = sk; brow = rq[q]; }
        }
      }
      argmax_tree(best, brow);
      if (lane == 0) { wbest[warp] = best; wrow[warp] = brow; }
      __syncthreads();
      unsigned long long cb = lane < NW ? wbest[lane] : 0ull;
      int crow = lane < NW ? wrow[lane] : INT_MAX;
      argmax_tree(cb, crow);
      const int krow_cta = K / RT;
#pragma unroll
      for (int q = 0; q < RPT; ++q) {
        if (cb != 0ull && rq[q] == crow) {
#pragma unroll
          for (int c = 0; c < NBP; ++c) myrow[c] = a[q][c];
        }
        if (rq[q] == K) {
#pragma unroll
          for (int c = 0; c < NBP; ++c) mykrow[c] = a[q][c];
        }
      }
      __syncthreads();
      for (int idx = tid; idx < CL * NBP; idx += THREADS) {
        const int peer = idx / NBP, c = idx % NBP;
        RegPanelTables<T, NBP>* rt = cluster.map_shared_rank(&tab, peer);
        if (cb != 0ull) rt->vals[par][cta][c] = myrow[c];
        if (cta == krow_cta) rt->rowk[par][c] = mykrow[c];
      }
      if (tid < CL) {
        RegPanelTables<T, NBP>* rt = cluster.map_shared_rank(&tab, tid);
        rt->score[par][cta] = cb;
        rt->row[par][cta] = crow;
      }
      cluster.sync();
      unsigned long long gb = 0ull;
      int grow = INT_MAX;
      if (lane < CL) {
        gb = tab.score[par][lane];
        if (gb != 0ull) grow = tab.row[par][lane];
      }
      const int my_row = grow;
      argmax_tree(gb, grow);
      const bool nonzero = gb > 1ull;
      const int gw = __ffs(__ballot_sync(0xffffffffu, my_row == grow && grow != INT_MAX)) - 1;
      const bool have = nonzero && gw >= 0;
      const int piv = have ? grow : K;
      const T* urow = have ? tab.vals[par][gw] : tab.rowk[par];
      if (tid == 0) {
        pivrel[K] = piv;
        if (cta == 0) {
          ipiv_s[K] = (int)(row_base + piv + 1);
          if (!have) atomicMin(info, (int)(col_base + K + 1));
        }
      }
      const bool helper = cta == 0 && warp == 1;
      const bool swap_l = helper && piv != K && lane < kk;
      T t1 = Sc<T>::zero(), t2 = Sc<T>::zero();
      T* lk = nullptr;
      T* lp = nullptr;
      if (helper && lane < NBP) Ubuf[K][lane] = urow[lane];
      if (swap_l) {
        lk = cluster.map_shared_rank(Lbuf, K / RT) + (size_t)lane * RT + K % RT;
        lp = cluster.map_shared_rank(Lbuf, piv / RT) + (size_t)lane * RT + piv % RT;
        t1 = *lk; t2 = *lp;
      }
      const T inv_pv = have ? sc_fast_recip<T>(urow[0]) : Sc<T>::zero();
#pragma unroll
      for (int q = 0; q < RPT; ++q) {
        if (piv != K && rq[q] == piv) {
#pragma unroll
          for (int c = 0; c < NBP; ++c) a[q][c] = tab.rowk[par][c];
        }
        const bool act = rq[q] > K && rq[q] < mrows;
        T l = Sc<T>::zero();
        if (act) {
          l = have ? Sc<T>::mul(a[q][0], inv_pv) : a[q][0];
          Lbuf[(size_t)kk * RT + q * THREADS + tid] = l;
          if (!have) l = Sc<T>::zero();
        }
#pragma unroll
        for (int j = 1; j < NBP; ++j) {
          T v = a[q][j];
          sc_fnma<T>(v, l, urow[j]);
          a[q][j - 1] = v;
        }
        a[q][NBP - 1] = Sc<T>::zero();
      }
      if (swap_l) { *lk = t2; *lp = t1; }
    }
  };

  run_phase(0, NBP);
  cluster.sync();   // every multiplier and every interchange of phase 0 has landed in its column buffer

  // ---- between the phases ------------------------------------------------------------------------------------------------
  // source row of a panel position for the columns that have not been touched yet: undo the interchanges in reverse order
  auto source_row = [&](int pos, int k_first, int k_last) {
    for (int k = k_last; k >= k_first; --k) {
      const int pk = pivrel[k];
      if (pos == k) pos = pk; else if (pos == pk) pos = k;
    }
    return pos;
  };
#pragma unroll
  for (int q = 0; q < RPT; ++q) {
    const int src = rq[q] < mrows ? source_row(rq[q], 0, NBP - 1) : 0;
#pragma unroll
    for (int c = 0; c < NBP; ++c) a[q][c] = (rq[q] < mrows && c < nb1) ? A[src + (int64_t)(NBP + c) * lda] : Sc<T>::zero();
  }
  // unit-lower L11 (CTA 0's column buffer, rows 0 .. NBP-1) and the gathered top block, into this CTA's shared memory
  {
    const T* L0 = cluster.map_shared_rank(Lbuf, 0);
    for (int idx = tid; idx < NBP * NBP; idx += THREADS) {
      const int i = idx % NBP, j = idx / NBP;
      L11s[i][j] = i > j ? L0[(size_t)j * RT + i] : Sc<T>::zero();
    }
    for (int idx = tid; idx < NBP * NBP; idx += THREADS) {
      const int i = idx % NBP, c = idx / NBP;
      const int src = source_row(i, 0, NBP - 1);
      U12[i][c] = c < nb1 ? A[src + (int64_t)(NBP + c) * lda] : Sc<T>::zero();
    }
  }
  __syncthreads();
  if (warp == 0 && lane < NBP) {   // U12 := L11^-1 U12, one column per lane, rotated forward substitution (unit diagonal)
    T x[NBP];
#pragma unroll
    for (int i = 0; i < NBP; ++i) x[i] = U12[i][lane];
#pragma unroll 1
    for (int q = 0; q < NBP; ++q) {
      const T xq = x[0];
      U12[q][lane] = xq;
#pragma unroll
      for (int i = 1; i < NBP; ++i) {
        T v = x[i];
        if (q + i < NBP) sc_fnma<T>(v, L11s[q + i][q], xq);
        x[i - 1] = v;
      }
    }
  }
  __syncthreads();
  // rows below the first NBP: subtract L(r, 0:NBP) U12 using this thread's own multipliers; then flush phase 0
#pragma unroll
  for (int q = 0; q < RPT; ++q) {
    if (rq[q] >= NBP && rq[q] < mrows) {
#pragma unroll 4
      for (int k = 0; k < NBP; ++k) {
        const T l = Lbuf[(size_t)k * RT + q * THREADS + tid];
#pragma unroll
        for (int c = 0; c < NBP; ++c) sc_fnma<T>(a[q][c], l, U12[k][c]);
      }
    }
    if (rq[q] < mrows) {
      const int ncol = min(rq[q], NBP);
      for (int c = 0; c < ncol; ++c) A[rq[q] + (int64_t)c * lda] = Lbuf[(size_t)c * RT + q * THREADS + tid];
    }
  }
  if (cta == 0) {
    for (int idx = tid; idx < NBP * NBP; idx += THREADS) {
      const int k = idx / NBP, j = idx % NBP;
      if (k + j < NBP) A[k + (int64_t)(k + j) * lda] = Ubuf[k][j];                 // U of phase 0's own columns
      if (j < nb1) A[k + (int64_t)(NBP + j) * lda] = U12[k][j];                      // rows 0 .. NBP-1 of phase 1's columns
    }
  }
  cluster.sync();   // every remote read of phase 0's column buffers is done before phase 1 overwrites them

  const int steps1 = nb1;   // mrows >= 2 NBP: every column of phase 1 has a pivot row
  run_phase(NBP, steps1);
  cluster.sync();
#pragma unroll
  for (int q = 0; q < RPT; ++q) {
    if (rq[q] < mrows) {
      const int ncol = max(0, min(rq[q] - NBP, steps1));   // columns NBP + c with NBP + c < r hold L(r, NBP + c)
      for (int c = 0; c < ncol; ++c) A[rq[q] + (int64_t)(NBP + c) * lda] = Lbuf[(size_t)c * RT + q * THREADS + tid];
    }
  }
  if (cta == 0) {
    for (int idx = tid; idx < steps1 * NBP; idx += THREADS) {
      const int k = idx / NBP, j = idx % NBP;
      if (k + j < nb1) A[(NBP + k) + (int64_t)(NBP + k + j) * lda] = Ubuf[NBP + k][j];
    }
    if (tid < NBP + steps1) ipiv[tid] = ipiv_s[tid];
  }
  cluster.sync();   // every CTA's flush of phase 0 is visible: phase 1's interchanges on the L part of columns 0 .. NBP-1
  if (cta == 0) {
    // destinations: NBP + t (t < steps1) and the pivot rows outside that range; 2 * steps1 destinations x NBP columns
    constexpr int PAIRS = (2 * NBP * NBP + THREADS - 1) / THREADS;
    T v[PAIRS];
    int dsts[PAIRS], cols[PAIRS];
#pragma unroll
    for (int e = 0; e < PAIRS; ++e) {
      const int pidx = tid + e * THREADS;
      const int d = pidx % (2 * NBP), c = pidx / (2 * NBP);
      int dst = -1;
      if (c < NBP) {
        if (d < steps1) dst = NBP + d;
        else if (d - NBP >= 0 && d - NBP < steps1 && pivrel[NBP + d - NBP] >= NBP + steps1) dst = pivrel[NBP + d - NBP];
      }
      int src = dst;
      if (dst >= 0) src = source_row(dst, NBP, NBP + steps1 - 1);
      dsts[e] = (dst >= 0 && src != dst) ? dst : -1;
      cols[e] = c;
      v[e] = dsts[e] >= 0 ? __ldcg(&A[src + (int64_t)c * lda]) : Sc<T>::zero();
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < PAIRS; ++e)
      if (dsts[e] >= 0) A[dsts[e] + (int64_t)cols[e] * lda] = v[e];
  }
  cluster.sync();   // nobody exits while a peer may still access its shared memory
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

// all three steps for ns <= 256 interchanges in ONE launch: a CTA owns whole columns, so reading every source of a column
// before writing any destination needs only a __syncthreads (the displaced list has at most ns entries)
template <typename T>
__global__ void __launch_bounds__(256) perm_apply_fused_kernel(int ns, int64_t ncols, const int* __restrict__ src_top, const int* __restrict__ count,
                                                               const int* __restrict__ dst, const int* __restrict__ src, T* __restrict__ A, int64_t lda) {
  const int tid = threadIdx.x;
  const int s_top = tid < ns ? src_top[tid] : -1;
  int d = -1, sd = 0;
  if (tid < *count) { d = dst[tid]; sd = src[tid]; }
  for (int64_t c = blockIdx.x; c < ncols; c += gridDim.x) {
    T* col = A + c * lda;
    T vtop = Sc<T>::zero(), vdisp = Sc<T>::zero();
    if (s_top >= 0) vtop = col[s_top];
    if (d >= 0) vdisp = col[sd];
    __syncthreads();
    if (s_top >= 0 && s_top != tid) col[tid] = vtop;
    if (d >= 0) col[d] = vdisp;
  }
}

// Up to 128 interchanges: no separate permutation pass at all.  Only the <= 2 ns rows {k} and {piv_k} change; every CTA
// finds the source of each of them itself by undoing the interchanges in reverse order (one thread per destination,
// ns steps of two compares on the pivot list in shared memory), then moves its columns: read every source, __syncthreads,
// write every destination.
constexpr int DIRECT_NS = 128;
template <typename T>
__global__ void __launch_bounds__(256) perm_apply_direct_kernel(const int* __restrict__ ipiv, int64_t k0, int ns, int mrows, int64_t ncols,
                                                                T* __restrict__ A, int64_t lda) {
  __shared__ int piv[DIRECT_NS];
  const int tid = threadIdx.x;
  if (tid < ns) {
    const int pv = ipiv[k0 + tid] - 1 - (int)k0;
    piv[tid] = (pv >= 0 && pv < mrows) ? pv : tid;
  }
  __syncthreads();
  // destination of this thread: top rows 0 .. ns-1 (threads 0 .. ns-1), far rows piv_t >= ns (threads ns .. 2 ns - 1)
  int dst = -1;
  if (tid < ns) dst = tid;
  else if (tid < 2 * ns && piv[tid - ns] >= ns) dst = piv[tid - ns];
  int src = dst;
  if (dst >= 0) {
    for (int k = ns - 1; k >= 0; --k) {
      const int pk = piv[k];
      if (src == k) src = pk; else if (src == pk) src = k;
    }
  }
  const bool moves = dst >= 0 && src != dst;
  for (int64_t c = blockIdx.x; c < ncols; c += gridDim.x) {
    T* col = A + c * lda;
    T v = Sc<T>::zero();
    if (moves) v = col[src];
    __syncthreads();
    if (moves) col[dst] = v;
  }
}

// the permutation of the interchanges ipiv[k0 .. k0+ns) (rows k0 .. m) into cx's lists
inline int build_perm(const GetrfProblem& p, const GetrfCtx& cx, int64_t k0, int64_t ns, cudaStream_t s) {
  if (ns <= DIRECT_NS) return 0;   // perm_apply_direct_kernel needs no lists
  const int64_t mrows = p.m - k0;
  if ((size_t)mrows * sizeof(int) + 2048 <= cx.max_dyn_smem) {
    perm_build_kernel<<<1, 1024, (size_t)mrows * sizeof(int), s>>>(p.dipiv, k0, (int)ns, (int)mrows, cx.src_top, cx.disp_dst, cx.disp_src, cx.disp_count);
  } else {
    perm_build_global_kernel<<<1, 32, 0, s>>>(p.dipiv, k0, (int)ns, (int)mrows, cx.idx_global, cx.src_top, cx.disp_dst, cx.disp_src, cx.disp_count);
  }
  count_launch();
  return (int)cudaGetLastError();
}
// apply the permutation in cx's lists to columns [c0, c0 + ncols) of A
template <typename T>
int apply_perm(const GetrfProblem& p, const GetrfCtx& cx, int64_t k0, int64_t ns, int64_t c0, int64_t ncols, cudaStream_t s) {
  if (ns <= 0 || ncols <= 0) return 0;
  T* A = (T*)p.A + k0;   // rows relative to k0
  if (ns <= DIRECT_NS) {
    const unsigned grid = (unsigned)std::min<int64_t>(ncols, (int64_t)cx.sms * 8);
    perm_apply_direct_kernel<T><<<grid, 256, 0, s>>>(p.dipiv, k0, (int)ns, (int)std::min<int64_t>(p.m - k0, INT_MAX), ncols, A + c0 * p.lda, p.lda);
    count_launch();
    return (int)cudaGetLastError();
  }
  if (ns <= 256) {
    const unsigned grid = (unsigned)std::min<int64_t>(ncols, (int64_t)cx.sms * 8);
    perm_apply_fused_kernel<T><<<grid, 256, 0, s>>>((int)ns, ncols, cx.src_top, cx.disp_count, cx.disp_dst, cx.disp_src, A + c0 * p.lda, p.lda);
    count_launch();
    return (int)cudaGetLastError();
  }
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
// apply the interchanges ipiv[k0 .. k0+ns) to columns [c0, c0 + ncols) of A (rows k0 .. m)
template <typename T>
int apply_pivots(const GetrfProblem& p, const GetrfCtx& cx, int64_t k0, int64_t ns, int64_t c0, int64_t ncols, cudaStream_t s) {
  if (ns <= 0 || ncols <= 0) return 0;
  B200_CUDA_TRY(build_perm(p, cx, k0, ns, s));
  return apply_perm<T>(p, cx, k0, ns, c0, ncols, s);
}

// register-panel shapes per type: up to ROWS_WIDE rows the leaf is NBP_WIDE columns wide with RPT_WIDE rows per thread,
// up to ROWS_TALL rows it is NBP_TALL x RPT_TALL (the register budget of 512 threads per CTA: 128 registers each)
constexpr int REG_THREADS = 512, REG_MAXCL = 16;
template <typename T> struct RegPanel {   // double, complex<float>: 8-byte scalars
  static constexpr int NBP_WIDE = 32, RPT_WIDE = 1, NBP_TALL = 16, RPT_TALL = 2;
  static constexpr int64_t ROWS_WIDE = (int64_t)REG_MAXCL * REG_THREADS * RPT_WIDE, ROWS_TALL = (int64_t)REG_MAXCL * REG_THREADS * RPT_TALL;
};
template <> struct RegPanel<float> {
  static constexpr int NBP_WIDE = 32, RPT_WIDE = 1, NBP_TALL = 32, RPT_TALL = 2;
  static constexpr int64_t ROWS_WIDE = (int64_t)REG_MAXCL * REG_THREADS * RPT_WIDE, ROWS_TALL = (int64_t)REG_MAXCL * REG_THREADS * RPT_TALL;
};
template <> struct RegPanel<double2> {
  static constexpr int NBP_WIDE = 16, RPT_WIDE = 1, NBP_TALL = 8, RPT_TALL = 2;
  static constexpr int64_t ROWS_WIDE = (int64_t)REG_MAXCL * REG_THREADS * RPT_WIDE, ROWS_TALL = (int64_t)REG_MAXCL * REG_THREADS * RPT_TALL;
};
// leaf width of the panel recursion for a panel of `mrows` rows
template <typename T>
int reg_leaf_width(int64_t mrows) {
  static const int mode = [] { const char* e = getenv("B200BLAS_GETF2"); return !e ? 0 : (e[0] == 's' ? 1 : (e[0] == 'c' ? 2 : 0)); }();
  if (mode != 0 || mrows > RegPanel<T>::ROWS_TALL) return Leaf<T>::NB;
  const int w = mrows <= RegPanel<T>::ROWS_WIDE ? RegPanel<T>::NBP_WIDE : RegPanel<T>::NBP_TALL;
  // two chained phases in one launch (getf2_reg2_kernel) double the leaf width.  Measured on B200 (profiles/bench_r02/
  // level3_pass11_chained_leaf.txt): sgetrf 8192 35.9 -> 32.2 ms, dgetrf 8192 36.2 -> 36.8 ms (the 8-byte types run the chained
  // kernel at 128 registers with spills in the column loop, 181 us against 2 x 64 us + the folded launches), so the default
  // is on for float only; B200BLAS_GETF2_CHAIN=1 / 0 forces it on / off for every type.
  static const int chain_env = [] { const char* e = getenv("B200BLAS_GETF2_CHAIN"); return !e ? -1 : (e[0] == '0' ? 0 : 1); }();
  const bool chain = chain_env >= 0 ? chain_env == 1 : sizeof(T) == 4;
  return (chain && mrows >= 2 * w) ? 2 * w : w;
}

template <typename T, int NBP, int RPT>
int launch_reg_panel(int64_t mrows, int nb, T* A, int64_t lda, int* ipiv, int64_t j0, int* info, cudaStream_t s) {
  // up to NBP columns: one phase; up to 2 NBP (only offered when the panel has at least 2 NBP rows): two chained phases
  auto kern = nb > NBP ? getf2_reg2_kernel<T, NBP, RPT, REG_THREADS> : getf2_reg_kernel<T, NBP, RPT, REG_THREADS>;
  if (nb > NBP && (nb > 2 * NBP || mrows < 2 * NBP)) return (int)cudaErrorInvalidValue;
  {
    static std::atomic<uint64_t> done{0};
    int dev = 0;
    B200_CUDA_TRY(cudaGetDevice(&dev));
    if (!((done.load(std::memory_order_relaxed) >> (dev & 63)) & 1ull)) {
      B200_CUDA_TRY(cudaFuncSetAttribute(getf2_reg_kernel<T, NBP, RPT, REG_THREADS>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
      B200_CUDA_TRY(cudaFuncSetAttribute(getf2_reg2_kernel<T, NBP, RPT, REG_THREADS>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
      done.fetch_or(1ull << (dev & 63), std::memory_order_relaxed);
    }
  }
  int cl = (int)((mrows + (int64_t)REG_THREADS * RPT - 1) / ((int64_t)REG_THREADS * RPT));
  if (cl < 1) cl = 1;
  cudaLaunchConfig_t cfg = {};
  constexpr size_t lbuf_bytes = (size_t)NBP * RPT * REG_THREADS * sizeof(T);   // 128 KB in every configuration
  B200_SET_MAX_DYN_SMEM_ONCE((getf2_reg_kernel<T, NBP, RPT, REG_THREADS>), lbuf_bytes);
  B200_SET_MAX_DYN_SMEM_ONCE((getf2_reg2_kernel<T, NBP, RPT, REG_THREADS>), lbuf_bytes);
  cfg.gridDim = dim3(cl); cfg.blockDim = dim3(REG_THREADS); cfg.dynamicSmemBytes = lbuf_bytes; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  B200_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, mrows, nb, A, lda, ipiv, j0, info, j0));
  count_launch();
  return 0;
}

template <typename T>
int launch_panel(const GetrfProblem& p, const GetrfCtx& cx, int64_t j0, int64_t nc, cudaStream_t s) {
  constexpr int NBP = Leaf<T>::NB;
  const int64_t mrows = p.m - j0;
  // B200BLAS_GETF2: "reg" (default) register-resident cluster panel; "slab" round 1's cooperative shared-memory slab kernel;
  // "cluster" the slab kernel on one cluster (forced variants for the sweeps)
  static const int mode = [] { const char* e = getenv("B200BLAS_GETF2"); return !e ? 0 : (e[0] == 's' ? 1 : (e[0] == 'c' ? 2 : 0)); }();
  if (mode == 0 && mrows <= RegPanel<T>::ROWS_TALL && nc <= reg_leaf_width<T>(mrows)) {
    T* Ap = (T*)p.A + j0 + j0 * p.lda;
    if (mrows <= RegPanel<T>::ROWS_WIDE)
      return launch_reg_panel<T, RegPanel<T>::NBP_WIDE, RegPanel<T>::RPT_WIDE>(mrows, (int)nc, Ap, p.lda, p.dipiv + j0, j0, p.dinfo, s);
    return launch_reg_panel<T, RegPanel<T>::NBP_TALL, RegPanel<T>::RPT_TALL>(mrows, (int)nc, Ap, p.lda, p.dipiv + j0, j0, p.dinfo, s);
  }
  const bool use_cluster = mode == 2;
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
  const int NBP = reg_leaf_width<T>(p.m - j0);
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

// B := inv(L11) * B for the unit-lower jb x jb block at the top of panel j0, B = jb x ncols at column c0 (PartialPivLU.h:490):
// one launch of the shared-memory block solve (tri.cu), or -- B200BLAS_TRSM=inv -- one product on the tensor-pipe kernels
// against the block's inverse (Vinv, from launch_trtri_diag) plus a copy back
template <typename T>
int panel_row_solve(const GetrfProblem& p, int64_t j0, int64_t jb, int64_t c0, int64_t ncols, const void* Vinv, void* Xtmp, cudaStream_t s) {
  if (ncols <= 0) return 0;
  T* A = (T*)p.A;
  TriProblem t;
  t.type = p.type; t.left = 1; t.uplo = UPLO_LOWER; t.op = OP_N; t.unit = 1; t.m = jb; t.n = ncols;
  t.alpha[0] = 1.0; t.alpha[1] = 0.0;
  t.A = A + j0 + j0 * p.lda; t.lda = p.lda; t.B = A + j0 + c0 * p.lda; t.ldb = p.lda;
  if (trsm_inverse_forced()) { t.Vinv = Vinv; t.Xtmp = Xtmp; }   // default: the shared-memory block solve of tri.cu (one launch, in place)
  return launch_trsm(t, s);
}
// A[r0.., c0..c0+ncols) -= A[r0.., j0..j0+jb) * A[j0..j0+jb, c0..c0+ncols)   (PartialPivLU.h:492)
template <typename T>
int panel_trailing_update(const GetrfProblem& p, int64_t j0, int64_t jb, int64_t c0, int64_t ncols, cudaStream_t s) {
  const int64_t r0 = j0 + jb, mlow = p.m - r0;
  if (mlow <= 0 || ncols <= 0) return 0;
  T* A = (T*)p.A;
  GemmProblem g;
  g.type = p.type; g.opa = OP_N; g.opb = OP_N; g.m = mlow; g.n = ncols; g.k = jb;
  g.alpha[0] = -1.0; g.alpha[1] = 0.0; g.beta[0] = 1.0; g.beta[1] = 0.0;
  g.A = A + r0 + j0 * p.lda; g.lda = p.lda;
  g.B = A + j0 + c0 * p.lda; g.ldb = p.lda;
  g.C = A + r0 + c0 * p.lda; g.ldc = p.lda;
  return run_gemm_device(g, s, B200BLAS_AUTO);
}

// Right-looking blocked LU with one step of look-ahead -- the blocked loop of PartialPivLU.h:426-496 with blockSize = NBO:
//   panel chain (stream sp): factor the panel (recursion over register-resident leaves) -> invert its L11 -> interchanges,
//                            row solve and update of the NEXT block column only -> factor the next panel ...
//   bulk (the caller's stream): interchanges left of the panel and right of the next block column, row solve and the large
//                            update A22 -= A21 A12 there,
// so that the latency-bound panel of step j+1 runs underneath the tensor-pipe update of step j.  cx / cx2 hold the
// permutation lists of the two streams.
template <typename T>
int getrf_blocked(const GetrfProblem& p, const GetrfCtx& cx, const GetrfCtx& cx2, void* Vinv0, void* Vinv1, void* Xtmp, void* Xtmp2, cudaStream_t s) {
  const int NBO = trsm_leaf_order(p.type);   // 128 (64 for complex double): one inverse block per panel
  const int64_t size = std::min(p.m, p.n);
  const bool look = lookahead_enabled() && size > 4 * NBO;
  cudaStream_t sp = s;
  LookAhead& la = t_look;
  if (look) {
    B200_CUDA_TRY(la.init());
    sp = la.sp;
    B200_CUDA_TRY(cudaEventRecord(la.e_in, s));
    B200_CUDA_TRY(cudaStreamWaitEvent(sp, la.e_in, 0));
  }
  bool rest_pending = false;
  for (int64_t j = 0; j < size; j += NBO) {
    const int64_t jb = std::min<int64_t>(NBO, size - j), r1 = j + jb;
    const int64_t nb2 = r1 < size ? std::min<int64_t>(NBO, size - r1) : 0;   // width of the next panel
    const int64_t c_rest = r1 + nb2, n_rest = p.n - c_rest;
    B200_CUDA_TRY(getrf_rec<T>(p, cx, j, jb, sp));
    // two inverse buffers, alternating: the bulk update of step j may still read its inverse while the panel chain of
    // step j+1 writes the next one (the chain only passes step j+1 after the bulk update of step j, see e_rest)
    void* Vinv = ((j / NBO) & 1) ? Vinv1 : Vinv0;
    const bool have_right = p.n > r1;
    if (have_right) {
      TriProblem t;
      t.type = p.type; t.left = 1; t.uplo = UPLO_LOWER; t.op = OP_N; t.unit = 1; t.m = jb; t.n = 1;
      t.A = (T*)p.A + j + j * p.lda; t.lda = p.lda;
      if (trsm_inverse_forced()) B200_CUDA_TRY(launch_trtri_diag(t, Vinv, sp));
    }
    if (look) B200_CUDA_TRY(cudaEventRecord(la.e_panel, sp));
    if (nb2 > 0) {   // next block column on the panel stream
      if (look && rest_pending) B200_CUDA_TRY(cudaStreamWaitEvent(sp, la.e_rest, 0));   // the previous bulk update wrote these columns
      B200_CUDA_TRY(build_perm(p, cx, j, jb, sp));
      B200_CUDA_TRY(apply_perm<T>(p, cx, j, jb, r1, nb2, sp));
      B200_CUDA_TRY(panel_row_solve<T>(p, j, jb, r1, nb2, Vinv, Xtmp, sp));
      B200_CUDA_TRY(panel_trailing_update<T>(p, j, jb, r1, nb2, sp));
    }
    if (j > 0 || n_rest > 0) {   // bulk: columns left of the panel and right of the next block column
      if (look) B200_CUDA_TRY(cudaStreamWaitEvent(s, la.e_panel, 0));
      B200_CUDA_TRY(build_perm(p, cx2, j, jb, s));
      B200_CUDA_TRY(apply_perm<T>(p, cx2, j, jb, 0, j, s));
      if (n_rest > 0) {
        B200_CUDA_TRY(apply_perm<T>(p, cx2, j, jb, c_rest, n_rest, s));
        B200_CUDA_TRY(panel_row_solve<T>(p, j, jb, c_rest, n_rest, Vinv, Xtmp2, s));
        B200_CUDA_TRY(panel_trailing_update<T>(p, j, jb, c_rest, n_rest, s));
      }
      if (look) { B200_CUDA_TRY(cudaEventRecord(la.e_rest, s)); rest_pending = true; }
    }
  }
  if (look) {
    B200_CUDA_TRY(cudaEventRecord(la.e_out, sp));
    B200_CUDA_TRY(cudaStreamWaitEvent(s, la.e_out, 0));
  }
  return 0;
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
  // default: right-looking blocked loop with look-ahead; B200BLAS_GETRF=rec keeps round 1's full recursion (forced variant)
  static const bool recursive = [] { const char* e = getenv("B200BLAS_GETRF"); return e && e[0] == 'r'; }();
  // workspace: two sets of permutation lists (3 * size + 16 ints, + m ints for the slow path; one set per stream), panel
  // scratch, gather buffer, the inverse of a panel's L11 and two scratch panels for the row solves
  const size_t ints_one = ((size_t)size * 3 + 16 + (size_t)p.m + 63) / 64 * 64;
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
  const size_t lb = (size_t)trsm_leaf_order(p.type);
  const size_t vinv_bytes = recursive ? 0 : (lb * lb * sizeof(T) + 255) / 256 * 256;
  const size_t xtmp_bytes = recursive ? 0 : (lb * (size_t)p.n * sizeof(T) + 255) / 256 * 256;
  const size_t total = 2 * ints_one * sizeof(int) + (ps + 255) / 256 * 256 + (cx.w_bytes + 255) / 256 * 256 + (gslab_bytes + 255) / 256 * 256 +
                       2 * vinv_bytes + 2 * xtmp_bytes;
  unsigned char* ws = nullptr;
  B200_CUDA_TRY(cudaMallocAsync((void**)&ws, total, s));
  auto lists = [&](GetrfCtx& c, int* ip) {
    c.src_top = ip; c.disp_dst = ip + size; c.disp_src = ip + 2 * size; c.disp_count = ip + 3 * size; c.idx_global = ip + 3 * size + 16;
  };
  lists(cx, (int*)ws);
  cx.panel_scratch = ws + 2 * ints_one * sizeof(int);
  cx.W = cx.panel_scratch + (ps + 255) / 256 * 256;
  unsigned char* after_w = (unsigned char*)cx.W + (cx.w_bytes + 255) / 256 * 256;
  if (gslab_bytes) cx.gslab = after_w;
  unsigned char* vinv = after_w + (gslab_bytes + 255) / 256 * 256;
  GetrfCtx cx2 = cx;   // the bulk stream's permutation lists (interchanges of <= 256 rows never use W)
  lists(cx2, (int*)ws + ints_one);
  int e;
  if (!recursive) {
    e = getrf_blocked<T>(p, cx, cx2, vinv, vinv + vinv_bytes, vinv + 2 * vinv_bytes, vinv + 2 * vinv_bytes + xtmp_bytes, s);
    if (e) cudaDeviceSynchronize();   // the panel stream may still be using the workspace
  } else {
    e = getrf_rec<T>(p, cx, 0, size, s);
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
  }
  cudaFreeAsync(ws, s);
  return e;
}

}  // namespace

// *p.dinfo must hold INT_MAX on entry; on exit it is the smallest failing 1-based index, or still INT_MAX
int launch_potrf(const PotrfProblem& p, cudaStream_t s) {
  if (p.n <= 0) return 0;
  // default: right-looking blocked loop with look-ahead; B200BLAS_POTRF=rec keeps round 1's recursion (forced variant for the sweeps)
  static const bool recursive = [] { const char* e = getenv("B200BLAS_POTRF"); return e && e[0] == 'r'; }();
  if (!recursive) {
    note_variant("potrf_blocked_lookahead_cta128+leaf+syrk");
    switch (p.type) {
      case TY_S: return potrf_blocked<float>(p, s);
      case TY_D: return potrf_blocked<double>(p, s);
      case TY_C: return potrf_blocked<float2>(p, s);
      default: return potrf_blocked<double2>(p, s);
    }
  }
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
  note_variant("getrf_blocked_lookahead_regpanel+inv+gemm");
  switch (p.type) {
    case TY_S: return getrf_typed<float>(p, s);
    case TY_D: return getrf_typed<double>(p, s);
    case TY_C: return getrf_typed<float2>(p, s);
    default: return getrf_typed<double2>(p, s);
  }
}

}  // namespace b200
