// eigen_b200/csrc/common.cuh -- shared declarations of libb200blas (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <stdint.h>
#include <stdio.h>

namespace b200 {

enum Op : int { OP_N = 0, OP_T = 1, OP_C = 2, OP_INVALID = 0xff };  // blas/common.h:24-42
enum Type : int { TY_S = 0, TY_D = 1, TY_C = 2, TY_Z = 3 };

// One product C = alpha*op(A)*op(B) + beta*C on device memory; leading dimensions in elements of the scalar type.
struct GemmProblem {
  int type;
  int opa, opb;
  int64_t m, n, k;
  double alpha[2], beta[2];  // widened on the host; kernels narrow to their scalar type
  const void* A; int64_t lda;
  const void* B; int64_t ldb;
  void* C; int64_t ldc;
  // rank-k updates (?syrk_/?herk_, blas/level3_impl.h:357-433,564-627) run on the same kernels with a triangular mask:
  int uplo = 0;   // 0: whole m x n window; UPLO_UPPER / UPLO_LOWER: only that triangle of C is read and written
  int herm = 0;   // 1: Hermitian update -- the imaginary part of the diagonal is stored as exactly zero
};
enum : int { UPLO_FULL = 0, UPLO_UPPER = 1, UPLO_LOWER = 2 };
// element (i, j) belongs to the referenced triangle
__host__ __device__ __forceinline__ bool in_triangle(int uplo, int64_t i, int64_t j) {
  return uplo == UPLO_FULL || (uplo == UPLO_UPPER ? i <= j : i >= j);
}
// the tile [i0, i1) x [j0, j1) has no element in the referenced triangle
__host__ __device__ __forceinline__ bool tile_outside(int uplo, int64_t i0, int64_t i1, int64_t j0, int64_t j1) {
  return uplo == UPLO_UPPER ? (i0 > j1 - 1) : uplo == UPLO_LOWER ? (i1 - 1 < j0) : false;
}

// B := alpha * inv(op(A)) * B / alpha * op(A) * B (left) or B := alpha * B * inv(op(A)) / alpha * B * op(A) (right); A triangular
// of order m (left) or n (right), B is m x n (?trsm_/?trmm_, blas/level3_impl.h:78-284).  Device pointers.
struct TriProblem {
  int type;
  int left;          // 1: side 'L', 0: side 'R'
  int uplo;          // UPLO_UPPER / UPLO_LOWER: the stored triangle of A
  int op;            // OP_N / OP_T / OP_C applied to A
  int unit;          // 1: unit diagonal, the stored diagonal is not referenced
  int64_t m, n;
  double alpha[2];
  const void* A; int64_t lda;
  void* B; int64_t ldb;
  // inverses of the diagonal leaf blocks (leaf order x leaf order each, ld = leaf order) and a
  // scratch panel for the out-of-place leaf products; null = substitution leaves
  const void* Vinv = nullptr; void* Xtmp = nullptr;
};
// C := alpha * A * B + beta * C (left) or alpha * B * A + beta * C (right); A symmetric (herm = 0) or Hermitian (herm = 1),
// only its `uplo` triangle is referenced (?symm_/?hemm_, blas/level3_impl.h:287-355,505-562).  Device pointers.
struct SymmProblem {
  int type;
  int left, uplo, herm;
  int64_t m, n;
  double alpha[2], beta[2];
  const void* A; int64_t lda;
  const void* B; int64_t ldb;
  void* C; int64_t ldc;
};

// In-place Cholesky A = L L^H / U^H U (?potrf_, lapack/cholesky.cpp:14-38) and LU with partial pivoting P A = L U (?getrf_,
// lapack/lu.cpp:14-42) on device memory.  *dinfo must hold INT_MAX on entry; the kernels atomicMin the first failing
// 1-based index into it.  dipiv receives min(m, n) 1-based row numbers.
struct PotrfProblem { int type; int uplo; int64_t n; void* A; int64_t lda; int* dinfo; };
struct GetrfProblem { int type; int64_t m, n; void* A; int64_t lda; int* dipiv; int* dinfo; };

static inline int type_bytes(int t) { return t == TY_S ? 4 : t == TY_Z ? 16 : 8; }

// kernel launchers (one translation unit each); return cudaError_t as int, bump the launch counter themselves
int launch_simt(const GemmProblem& p, cudaStream_t s);
int launch_dmma(const GemmProblem& p, cudaStream_t s);      // type D / Z
int launch_tf32x3(const GemmProblem& p, cudaStream_t s, void* workspace, size_t workspace_bytes);  // type S / C
size_t tf32x3_workspace_bytes(const GemmProblem& p);
bool dmma_supported(const GemmProblem& p);
bool tf32x3_supported(const GemmProblem& p);
double pipe_peak(int pipe, int millis);
// variant choice + launch of one product on device memory (host.cu); used by the composite level-3 routines
int run_gemm_device(const GemmProblem& p, cudaStream_t s, int variant);
int launch_trsm(const TriProblem& p, cudaStream_t s);   // tri.cu
int launch_trmm(const TriProblem& p, cudaStream_t s);
int trsm_leaf_order(int type);                                            // order of the diagonal blocks TriProblem::Vinv holds
int launch_trtri_diag(const TriProblem& p, void* V, cudaStream_t s);      // fills such a Vinv for p's triangle
bool trsm_substitution_forced();                                          // B200BLAS_TRSM=subst
bool trsm_inverse_forced();                                               // B200BLAS_TRSM=inv
size_t symm_workspace_bytes(const SymmProblem& p);
int launch_symm(const SymmProblem& p, cudaStream_t s, void* workspace);
int launch_potrf(const PotrfProblem& p, cudaStream_t s);   // lapack.cu
int launch_getrf(const GetrfProblem& p, cudaStream_t s);

// multi.cu: the multi-device partitioner behind ?gemm_ / b200blas_gemm_dev.  multi_wanted: N > 1 devices are enabled and
// the product is large enough; multi_gemm returns a cudaError_t as int.  host_origin: operands are host pointers
// (the call is synchronous); otherwise they live on the current device and the call is asynchronous on `stream`.
bool multi_wanted(const GemmProblem& p);
int multi_gemm(const GemmProblem& p, bool host_origin, cudaStream_t stream, uint64_t* h2d_bytes, uint64_t* d2h_bytes);
void multi_release();

void count_launch(int n = 1);
void note_variant(const char* name);

#define B200_CUDA_TRY(expr)                                  \
  do {                                                       \
    const int _e = (int)(expr);                              \
    if (_e != 0) return _e;                                  \
  } while (0)

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a driver call of tens of microseconds: inside a chain of
// thousands of small dependent launches (?trsm_ / ?potrf_ / ?getrf_) it would leave the GPU idle between kernels.
// The attribute is per device, so each call site remembers the devices it has configured (one static mask per
// template instantiation).
#define B200_SET_MAX_DYN_SMEM_ONCE(kernel, bytes)                                                              \
  do {                                                                                                          \
    static std::atomic<uint64_t> _done{0};                                                                      \
    int _dev = 0;                                                                                               \
    B200_CUDA_TRY(cudaGetDevice(&_dev));                                                                        \
    if (!((_done.load(std::memory_order_relaxed) >> (_dev & 63)) & 1ull)) {                                     \
      B200_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)));   \
      _done.fetch_or(1ull << (_dev & 63), std::memory_order_relaxed);                                           \
    }                                                                                                           \
  } while (0)

// ---- device helpers ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// cp.async with zero fill: copies src_bytes (<= BYTES) and zero-fills the rest of the BYTES-wide destination
template <int BYTES>
__device__ __forceinline__ void cp_async_zfill(void* smem_dst, const void* gmem_src, int src_bytes) {
  static_assert(BYTES == 4 || BYTES == 8 || BYTES == 16, "cp.async size");
  if constexpr (BYTES == 16)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(src_bytes));
  else
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2, %3;\n" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "n"(BYTES), "r"(src_bytes));
}
// plain cp.async of BYTES bytes (no source-size operand: ptxas turns a variable source size into ~6 extra instructions)
template <int BYTES>
__device__ __forceinline__ void cp_async_full(void* smem_dst, const void* gmem_src) {
  static_assert(BYTES == 4 || BYTES == 8 || BYTES == 16, "cp.async size");
  if constexpr (BYTES == 16)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(smem_dst)), "l"(gmem_src));
  else
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;\n" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "n"(BYTES));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

}  // namespace b200
