// eigen_b200/csrc/host.cu -- the C-ABI host layer of libb200blas.so.
//
// Mirrors the reference's BLAS seam for the dense product:
//   EIGEN_BLAS_FUNC(gemm)            blas/level3_impl.h:12-76   (argument checks, info codes, quick returns, beta)
//   OP()                             blas/common.h:39-42        (case-insensitive N/T/C)
//   xerbla_                          blas/xerbla.cpp:15-19      (weak, overridable)
//   parallelize_gemm under EIGEN_USE_BLAS is a single call (Parallelizer.h:88-97): one ?gemm_ per product.
// Host operands are staged through CUDA streams: A is uploaded once, B and C travel in column slabs so that the
// upload of slab j+1, the product on slab j and the download of slab j-1 overlap (the GPU analogue of the kc/nc
// panel streaming of GeneralMatrixMatrix.h:155-198).  There is no CPU fallback.
#include <nvtx3/nvToolsExt.h>

#include <atomic>
#include <chrono>
#include <climits>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/b200blas.h"
#include "common.cuh"
#include "staging.cuh"

namespace b200 {

static std::atomic<uint64_t> g_launches{0};
static std::atomic<int> g_forced_variant{-1};
static thread_local char t_variant[64] = "";
static thread_local char t_error[256] = "";
static thread_local uint64_t t_h2d = 0, t_d2h = 0;

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }
void note_variant(const char* name) {
  strncpy(t_variant, name, sizeof(t_variant) - 1);
  t_variant[sizeof(t_variant) - 1] = 0;
}
static int fail(int cuda_err) {
  if (cuda_err != 0) {
    snprintf(t_error, sizeof t_error, "%s", cudaGetErrorString((cudaError_t)cuda_err));
    cudaGetLastError();  // clear sticky-less errors
  }
  return cuda_err;
}

// B200BLAS_LOG=1: one line per F77 product call on stderr (shape, operand residency, kernel variant, wall ms)
static bool log_enabled() {
  static const bool v = [] { const char* e = getenv("B200BLAS_LOG"); return e && e[0] && e[0] != '0'; }();
  return v;
}
// B200BLAS_NVTX=1: one NVTX range per entry point (header-only nvtx3: no link dependency, a no-op unless a tool such as
// Nsight Systems is attached) -- SURVEY section 5, tracing.
static bool nvtx_enabled() {
  static const bool v = [] { const char* e = getenv("B200BLAS_NVTX"); return e && e[0] && e[0] != '0'; }();
  return v;
}
struct NvtxRange {
  bool on;
  explicit NvtxRange(const char* name) : on(nvtx_enabled()) { if (on) nvtxRangePushA(name); }
  ~NvtxRange() { if (on) nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};
static double wall_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
static void log_call(const char* name, char ta, char tb, int64_t m, int64_t n, int64_t k, const char* where, double ms, int err) {
  fprintf(stderr, "[b200blas] %.6s %c%c m=%lld n=%lld k=%lld operands=%s devices=%d variant=%s %.3f ms h2d=%llu d2h=%llu%s\n", name, ta, tb,
          (long long)m, (long long)n, (long long)k, where, b200blas_get_devices(), t_variant, ms, (unsigned long long)t_h2d,
          (unsigned long long)t_d2h, err ? " FAILED" : "");
}

static int op_of(char x) {
  return (x == 'N' || x == 'n') ? OP_N : (x == 'T' || x == 't') ? OP_T : (x == 'C' || x == 'c') ? OP_C : OP_INVALID;
}

static int forced_variant() {
  int v = g_forced_variant.load();
  if (v >= 0) return v;
  static int env_v = [] {
    const char* e = getenv("B200BLAS_VARIANT");
    if (!e) return (int)B200BLAS_AUTO;
    if (!strcmp(e, "simt")) return (int)B200BLAS_SIMT;
    if (!strcmp(e, "dmma")) return (int)B200BLAS_DMMA;
    if (!strcmp(e, "tf32x3")) return (int)B200BLAS_TF32X3;
    return (int)B200BLAS_AUTO;
  }();
  return env_v;
}

// Variant choice.  Tensor tiles are 128x128 (DMMA) / 128x256 (tcgen05); below about one tile of work, or when the
// operands cannot be fed to the tensor variant, the SIMT kernel wins (thresholds: profiles/variant_sweep_r01.md).
static int choose_variant(const GemmProblem& p, int requested) {
  int v = requested != B200BLAS_AUTO ? requested : forced_variant();
  const bool dz = (p.type == TY_D || p.type == TY_Z);
  if (v == B200BLAS_DMMA && !(dz && dmma_supported(p))) v = B200BLAS_AUTO;
  if (v == B200BLAS_TF32X3 && !(!dz && tf32x3_supported(p))) v = B200BLAS_AUTO;
  if (v != B200BLAS_AUTO) return v;
  if (p.k == 0) return B200BLAS_SIMT;
  const double work = (double)p.m * (double)p.n * (double)p.k;
  if (dz) {
    if (dmma_supported(p) && p.m * p.n >= 64 * 64 && p.k >= 8 && work >= 64.0 * 64.0 * 64.0) return B200BLAS_DMMA;
    return B200BLAS_SIMT;
  }
  if (tf32x3_supported(p) && p.m >= 128 && p.n >= 128 && p.k >= 32 && work >= 256.0 * 256.0 * 256.0)
    return B200BLAS_TF32X3;
  return B200BLAS_SIMT;
}

// The pack workspace comes from the stream-ordered allocator; without a release threshold the pool hands its
// memory back to the driver at every synchronisation and the next call pays for cudaMalloc again.
static void keep_pool_memory() {
  static thread_local int done_dev = -1;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev == done_dev) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    uint64_t thr = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  }
  done_dev = dev;
}

static int run_device(const GemmProblem& p, cudaStream_t s, int variant);
int run_gemm_device(const GemmProblem& p, cudaStream_t s, int variant) { return run_device(p, s, variant); }

static int run_device(const GemmProblem& p, cudaStream_t s, int variant) {
  const int v = choose_variant(p, variant);
  if (v == B200BLAS_DMMA) return fail(launch_dmma(p, s));
  if (v == B200BLAS_TF32X3) {
    const size_t ws = tf32x3_workspace_bytes(p);
    void* w = nullptr;
    keep_pool_memory();
    if (ws) { const int e = (int)cudaMallocAsync(&w, ws, s); if (e) return fail(e); }
    const int e = launch_tf32x3(p, s, w, ws);
    if (w) cudaFreeAsync(w, s);
    return fail(e);
  }
  return fail(launch_simt(p, s));
}

// ---- argument checks of blas/level3_impl.h:47-57 ---------------------------------------------------------
static int check_args(int opa, int opb, int64_t m, int64_t n, int64_t k, int64_t lda, int64_t ldb, int64_t ldc) {
  if (opa == OP_INVALID) return 1;
  if (opb == OP_INVALID) return 2;
  if (m < 0) return 3;
  if (n < 0) return 4;
  if (k < 0) return 5;
  if (lda < std::max<int64_t>(1, opa == OP_N ? m : k)) return 8;
  if (ldb < std::max<int64_t>(1, opb == OP_N ? k : n)) return 10;
  if (ldc < std::max<int64_t>(1, m)) return 13;
  return 0;
}
static const char* k_names[4] = {"SGEMM ", "DGEMM ", "CGEMM ", "ZGEMM "};

static void load_scalar(int type, const void* p, double out[2]) {
  switch (type) {
    case TY_S: out[0] = *(const float*)p; out[1] = 0; break;
    case TY_D: out[0] = *(const double*)p; out[1] = 0; break;
    case TY_C: out[0] = ((const float*)p)[0]; out[1] = ((const float*)p)[1]; break;
    default: out[0] = ((const double*)p)[0]; out[1] = ((const double*)p)[1]; break;
  }
}



// Host-operand product: stage, multiply, return.  Returns a cudaError_t (0 = ok).
static int run_host(int type, int opa, int opb, int64_t m, int64_t n, int64_t k, const double alpha[2], const void* a,
                    int64_t lda, const void* b, int64_t ldb, const double beta[2], void* c, int64_t ldc) {
  StageLease lease;
  if (!lease.ok()) return (int)cudaErrorInitializationError;
  Staging& st = lease.st();
  { const int e = st.init(); if (e) return e; }
  const size_t es = (size_t)type_bytes(type);
  const bool beta_zero = (beta[0] == 0.0 && beta[1] == 0.0);
  const bool have_product = k > 0;
  // device images: column-major, leading dimension = rows rounded up to 16 bytes * 2 (keeps every column 16-byte aligned)
  const int64_t ra = (opa == OP_N) ? m : k, ca = (opa == OP_N) ? k : m;
  const int64_t rb = (opb == OP_N) ? k : n, cb = (opb == OP_N) ? n : k;
  const int64_t q = 32 / (int64_t)es > 0 ? 32 / (int64_t)es : 1;
  const int64_t dlda = round_up(std::max<int64_t>(ra, 1), q), dldb = round_up(std::max<int64_t>(rb, 1), q),
                dldc = round_up(m, q);
  if (have_product) {
    { const int e = st.reserve(0, (size_t)dlda * (size_t)std::max<int64_t>(ca, 1) * es); if (e) return e; }
    { const int e = st.reserve(1, (size_t)dldb * (size_t)std::max<int64_t>(cb, 1) * es); if (e) return e; }
  }
  { const int e = st.reserve(2, (size_t)dldc * (size_t)n * es); if (e) return e; }
  char* dA = (char*)st.dbuf[0];
  char* dB = (char*)st.dbuf[1];
  char* dC = (char*)st.dbuf[2];
  t_h2d = t_d2h = 0;
  // pageable operands of at least a few MiB go through the pinned ring; pinned / small ones are copied directly
  const size_t ring_min = (size_t)4 << 20;
  const bool page_a = have_product && (size_t)ra * ca * es >= ring_min && is_pageable(a);
  const bool page_b = have_product && (size_t)rb * cb * es >= ring_min && is_pageable(b);
  const bool page_c = (size_t)m * n * es >= ring_min && is_pageable(c);
  auto upload = [&](bool paged, char* dst, size_t dpitch, const char* src, size_t spitch, size_t width, size_t ncols) -> int {
    if (paged) return st.ring_in.h2d(dst, dpitch, src, spitch, width, ncols, st.s_in);
    return (int)cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, ncols, cudaMemcpyHostToDevice, st.s_in);
  };

  // column slabs of C (and of op(B)): ~16 slabs, at least 512 columns wide, multiples of 256 columns
  const double total_bytes = (double)es * ((double)m * k + (double)k * n + 2.0 * m * n);
  int64_t slab = n;
  static const int slabs_env = [] { const char* e = getenv("B200BLAS_HOST_SLABS"); const int v = e ? atoi(e) : 0; return v > 0 ? v : 16; }();
  if (total_bytes > 32e6 && n >= 1024) slab = std::max<int64_t>(512, round_up((n + slabs_env - 1) / slabs_env, 256));
  int nslabs = (int)((n + slab - 1) / slab);
  if (nslabs > Staging::MAX_SLABS) { slab = round_up((n + Staging::MAX_SLABS - 1) / Staging::MAX_SLABS, 256); nslabs = (int)((n + slab - 1) / slab); }

  // A travels in k-chunks (contiguous column blocks for 'N', row blocks for 'T'/'C'); the first column slab of C
  // accumulates chunk by chunk (beta = 1 after the first) while later chunks are still on the bus, so the
  // start-up bubble is one chunk instead of all of A.  Later slabs see A resident and run at full k.
  std::thread downloader;
  std::mutex dl_mu;
  std::condition_variable dl_cv;
  int dl_ready = 0, dl_err = 0;
  bool dl_started = false, dl_abort = false;
  struct JoinGuard {   // never leave this function with a joinable thread (early error returns)
    std::thread& t; std::mutex& mu; std::condition_variable& cv; bool& abort;
    ~JoinGuard() { if (t.joinable()) { { std::lock_guard<std::mutex> l(mu); abort = true; } cv.notify_all(); t.join(); } }
  } join_guard{downloader, dl_mu, dl_cv, dl_abort};
  // B200BLAS_HOST_HEAD=simple: round 2's first schedule (8 chunks of A, head slabs complete their inputs before A1 travels)
  static const bool staircase = [] { const char* e = getenv("B200BLAS_HOST_HEAD"); return !(e && e[0] == 's'); }();
  int nac = 1;
  if (have_product && nslabs > 1 && k >= 8 * 512) nac = staircase ? Staging::MAX_ACHUNKS : 8;
  const int64_t kch = round_up((k + nac - 1) / nac, 256);
  auto upload_a_chunk = [&](int c) -> int {
    const int64_t k0 = (int64_t)c * kch, kc = std::min<int64_t>(kch, k - k0);
    if (kc <= 0) return 0;
    if (opa == OP_N) {
      B200_CUDA_TRY(upload(page_a, dA + (size_t)k0 * dlda * es, (size_t)dlda * es, (const char*)a + (size_t)k0 * lda * es,
                           (size_t)lda * es, (size_t)m * es, (size_t)kc));
    } else {
      B200_CUDA_TRY(upload(page_a, dA + (size_t)k0 * es, (size_t)dlda * es, (const char*)a + (size_t)k0 * es,
                           (size_t)lda * es, (size_t)kc * es, (size_t)m));
    }
    t_h2d += (uint64_t)m * kc * es;
    B200_CUDA_TRY(cudaEventRecord(st.ev_ac[c], st.s_in));
    return 0;
  };
  // inputs of column slab j: the slab of op(B) and (beta != 0) of C
  auto upload_slab_inputs = [&](int j) -> int {
    const int64_t j0 = (int64_t)j * slab, nj = std::min<int64_t>(slab, n - j0);
    if (have_product) {
      if (opb == OP_N) {
        B200_CUDA_TRY(upload(page_b, dB + (size_t)j0 * dldb * es, (size_t)dldb * es, (const char*)b + (size_t)j0 * ldb * es,
                             (size_t)ldb * es, (size_t)k * es, (size_t)nj));
      } else {
        B200_CUDA_TRY(upload(page_b, dB + (size_t)j0 * es, (size_t)dldb * es, (const char*)b + (size_t)j0 * es,
                             (size_t)ldb * es, (size_t)nj * es, (size_t)k));
      }
      t_h2d += (uint64_t)k * nj * es;
    }
    if (!beta_zero) {  // beta == 0: C is never read, so it is not uploaded either (blas/level3_impl.h:64)
      B200_CUDA_TRY(upload(page_c, dC + (size_t)j0 * dldc * es, (size_t)dldc * es, (const char*)c + (size_t)j0 * ldc * es,
                           (size_t)ldc * es, (size_t)m * es, (size_t)nj));
      t_h2d += (uint64_t)m * nj * es;
    }
    B200_CUDA_TRY(cudaEventRecord(st.ev_in[j], st.s_in));
    return (int)cudaStreamWaitEvent(st.s_comp, st.ev_in[j], 0);
  };
  auto slab_problem = [&](int j) {
    const int64_t j0 = (int64_t)j * slab, nj = std::min<int64_t>(slab, n - j0);
    GemmProblem p;
    p.type = type; p.opa = opa; p.opb = opb; p.m = m; p.n = nj; p.k = k;
    p.alpha[0] = alpha[0]; p.alpha[1] = alpha[1]; p.beta[0] = beta[0]; p.beta[1] = beta[1];
    p.A = dA; p.lda = dlda;
    p.B = (opb == OP_N) ? dB + (size_t)j0 * dldb * es : dB + (size_t)j0 * es; p.ldb = dldb;
    p.C = dC + (size_t)j0 * dldc * es; p.ldc = dldc;
    return p;
  };
  // slab j is complete on s_comp: send its m x nj window back (rows m..ldc-1 of the caller's C stay untouched)
  auto return_slab = [&](int j) -> int {
    const int64_t j0 = (int64_t)j * slab, nj = std::min<int64_t>(slab, n - j0);
    B200_CUDA_TRY(cudaEventRecord(st.ev_comp[j], st.s_comp));
    if (page_c) {
      // a downloader thread drains finished slabs through the pinned ring while this thread keeps uploading
      if (!dl_started) {
        dl_started = true;
        downloader = std::thread([&, dev = st.dev] {
          cudaSetDevice(dev);
          for (int jj = 0; jj < nslabs; ++jj) {
            {
              std::unique_lock<std::mutex> l(dl_mu);
              dl_cv.wait(l, [&] { return dl_ready > jj || dl_abort; });
              if (dl_abort) return;
            }
            const int64_t q0 = (int64_t)jj * slab, nq = std::min<int64_t>(slab, n - q0);
            int e = (int)cudaEventSynchronize(st.ev_comp[jj]);
            if (!e) e = st.ring_out.d2h((char*)c + (size_t)q0 * ldc * es, (size_t)ldc * es, dC + (size_t)q0 * dldc * es,
                                        (size_t)dldc * es, (size_t)m * es, (size_t)nq, st.s_out);
            if (e) { dl_err = e; return; }
          }
        });
      }
      { std::lock_guard<std::mutex> l(dl_mu); dl_ready = j + 1; }
      dl_cv.notify_all();
    } else {
      B200_CUDA_TRY(cudaStreamWaitEvent(st.s_out, st.ev_comp[j], 0));
      B200_CUDA_TRY(cudaMemcpy2DAsync((char*)c + (size_t)j0 * ldc * es, (size_t)ldc * es, dC + (size_t)j0 * dldc * es,
                                      (size_t)dldc * es, (size_t)m * es, (size_t)nj, cudaMemcpyDeviceToHost, st.s_out));
    }
    t_d2h += (uint64_t)m * nj * es;
    return 0;
  };
  // Head of the pipeline: all of A has to cross the bus before ANY slab can finish (2 GiB = ~40 ms of PCIe at 16384^3, against
  // 16 ms of arithmetic per slab), so the first `head` slabs accumulate chunk by chunk (beta = 1 after the first chunk)
  // while the later chunks of A are still in flight: enough arithmetic to cover the whole upload of A instead of one slab's
  // worth.  Uploads are issued in order of first use: B0 C0 A0 | B1 C1 | B2 C2 | B3 C3 | A1 | A2 ...  Later slabs see A resident
  // and run at full k.
  const int head = !have_product ? 0 : (nac > 1 ? std::min(nslabs, 4) : std::min(nslabs, 1));
  if (staircase && head > 1) {
    // Staircase: uploads alternate between chunks of A and the inputs of the head slabs -- S0 A0 | A1 S1 | A2 S2 | A3 S3 | A4 A5 ...
    // (S_h = B_h and C_h) -- and every arrival releases the products it completes the inputs of: chunk c against every slab
    // that has arrived, slab h against every chunk that has arrived.  The first product starts after one slab and one
    // (half-size) chunk instead of after four slabs and one chunk, and the arithmetic released per arrival grows with the bytes
    // still to come; the products accumulate into their slab in whatever order they are released (beta on the first one).
    const int nch = (int)((k + kch - 1) / kch);
    std::vector<int> done(head, 0);
    int a_arrived = 0, s_arrived = 0;
    auto product = [&](int cix, int h) -> int {
      const int64_t k0 = (int64_t)cix * kch, kc = std::min<int64_t>(kch, k - k0);
      GemmProblem q = slab_problem(h);
      q.k = kc;
      q.A = (opa == OP_N) ? dA + (size_t)k0 * dlda * es : dA + (size_t)k0 * es;
      q.B = (opb == OP_N) ? (const char*)q.B + (size_t)k0 * es : (const char*)q.B + (size_t)k0 * dldb * es;
      if (done[h] > 0) { q.beta[0] = 1.0; q.beta[1] = 0.0; }
      { const int e = run_device(q, st.s_comp, B200BLAS_AUTO); if (e) { cudaDeviceSynchronize(); return e; } }
      if (++done[h] == nch) return return_slab(h);
      return 0;
    };
    auto arrive_slab = [&]() -> int {
      const int h = s_arrived;
      { const int e = upload_slab_inputs(h); if (e) return e; }   // also makes s_comp wait for it
      ++s_arrived;
      for (int cix = 0; cix < a_arrived; ++cix) { const int e = product(cix, h); if (e) return e; }
      return 0;
    };
    { const int e = arrive_slab(); if (e) return e; }
    for (int cix = 0; cix < nch; ++cix) {
      { const int e = upload_a_chunk(cix); if (e) return e; }
      B200_CUDA_TRY(cudaStreamWaitEvent(st.s_comp, st.ev_ac[cix], 0));
      ++a_arrived;
      for (int h = 0; h < s_arrived; ++h) { const int e = product(cix, h); if (e) return e; }
      if (cix >= 1 && s_arrived < head) { const int e = arrive_slab(); if (e) return e; }
    }
    while (s_arrived < head) { const int e = arrive_slab(); if (e) return e; }
  } else
  for (int cix = 0; cix < nac && head > 0; ++cix) {
    const int64_t k0 = (int64_t)cix * kch, kc = std::min<int64_t>(kch, k - k0);
    if (kc <= 0) break;
    for (int h = 0; h < head; ++h) {
      if (cix == 0) { const int e = upload_slab_inputs(h); if (e) return e; }
      if (h == 0) { const int e = upload_a_chunk(cix); if (e) return e; }
      B200_CUDA_TRY(cudaStreamWaitEvent(st.s_comp, st.ev_ac[cix], 0));
      GemmProblem q = slab_problem(h);
      q.k = kc;
      q.A = (opa == OP_N) ? dA + (size_t)k0 * dlda * es : dA + (size_t)k0 * es;
      q.B = (opb == OP_N) ? (const char*)q.B + (size_t)k0 * es : (const char*)q.B + (size_t)k0 * dldb * es;
      if (cix > 0) { q.beta[0] = 1.0; q.beta[1] = 0.0; }
      { const int e = run_device(q, st.s_comp, B200BLAS_AUTO); if (e) { cudaDeviceSynchronize(); return e; } }
      if (k0 + kc >= k) { const int e = return_slab(h); if (e) return e; }
    }
  }
  for (int j = head; j < nslabs; ++j) {
    { const int e = upload_slab_inputs(j); if (e) return e; }
    const GemmProblem p = slab_problem(j);
    { const int e = run_device(p, st.s_comp, B200BLAS_AUTO); if (e) { cudaDeviceSynchronize(); return e; } }
    { const int e = return_slab(j); if (e) return e; }
  }
  if (dl_started) {
    downloader.join();
    dl_started = false;
    if (dl_err) return dl_err;
  }
  B200_CUDA_TRY(cudaStreamSynchronize(st.s_out));
  B200_CUDA_TRY(cudaStreamSynchronize(st.s_comp));
  B200_CUDA_TRY(cudaStreamSynchronize(st.s_in));
  return 0;
}

static int gemm_entry(int type, const char* ta, const char* tb, const int* pm, const int* pn, const int* pk,
                      const void* palpha, const void* a, const int* plda, const void* b, const int* pldb,
                      const void* pbeta, void* c, const int* pldc) {
  NvtxRange nvtx_range("b200blas ?gemm_");
  const int opa = op_of(*ta), opb = op_of(*tb);
  int info = check_args(opa, opb, *pm, *pn, *pk, *plda, *pldb, *pldc);
  if (info) return xerbla_(k_names[type], &info, 6);
  if (*pm == 0 || *pn == 0) return 0;
  double alpha[2], beta[2];
  load_scalar(type, palpha, alpha);
  load_scalar(type, pbeta, beta);
  t_error[0] = 0;
  int err;
  const bool dev_c = is_device_ptr(c);
  if (*pk > 0 && (is_device_ptr(a) != dev_c || is_device_ptr(b) != dev_c)) {
    // mixed residency (some operands in host memory, some on the device) is not a BLAS calling convention
    snprintf(t_error, sizeof t_error, "operands must be all host or all device pointers");
    info = -1;
    return xerbla_(k_names[type], &info, 6);
  }
  GemmProblem p;
  p.type = type; p.opa = opa; p.opb = opb; p.m = *pm; p.n = *pn; p.k = *pk;
  p.alpha[0] = alpha[0]; p.alpha[1] = alpha[1]; p.beta[0] = beta[0]; p.beta[1] = beta[1];
  p.A = a; p.lda = *plda; p.B = b; p.ldb = *pldb; p.C = c; p.ldc = *pldc;
  const double t0 = log_enabled() ? wall_ms() : 0.0;
  if (multi_wanted(p)) {
    // the parallel split happens INSIDE the product call, as in the reference (Parallelizer.h:85-157): multi.cu
    t_h2d = t_d2h = 0;
    err = fail(multi_gemm(p, !dev_c, nullptr, &t_h2d, &t_d2h));
    if (!err && dev_c) err = fail((int)cudaStreamSynchronize(nullptr));
  } else if (dev_c) {
    err = run_device(p, nullptr, B200BLAS_AUTO);
    if (!err) err = fail((int)cudaStreamSynchronize(nullptr));
  } else {
    err = fail(run_host(type, opa, opb, *pm, *pn, *pk, alpha, a, *plda, b, *pldb, beta, c, *pldc));
  }
  if (log_enabled()) log_call(k_names[type], *ta, *tb, *pm, *pn, *pk, dev_c ? "device" : "host", wall_ms() - t0, err);
  if (err) {
    // no CPU fallback by contract: CUDA failures surface through xerbla_ with the reserved info -1
    info = -1;
    return xerbla_(k_names[type], &info, 6);
  }
  return 0;
}


// ---- ?syrk_ / ?herk_: C.triangle = alpha*op(A)*op(A)^T|^H + beta*C.triangle -------------------------------
// Reference: blas/level3_impl.h:357-433 (syrk) and :564-627 (herk).  The product runs on the GEMM kernels with
// B := A and a triangular tile mask; only the referenced triangle of C is read or written.
static const char* k_syrk_names[4] = {"SSYRK ", "DSYRK ", "CSYRK ", "ZSYRK "};
static const char* k_herk_names[4] = {"", "", "CHERK ", "ZHERK "};

static int run_host_rankk(const GemmProblem& hp) {
  StageLease lease;
  if (!lease.ok()) return (int)cudaErrorInitializationError;
  Staging& st = lease.st();
  { const int e = st.init(); if (e) return e; }
  const size_t es = (size_t)type_bytes(hp.type);
  const int64_t n = hp.m, k = hp.k;
  const bool beta_zero = (hp.beta[0] == 0.0 && hp.beta[1] == 0.0);
  const int64_t ra = (hp.opa == OP_N) ? n : k, ca = (hp.opa == OP_N) ? k : n;
  const int64_t q = 32 / (int64_t)es > 0 ? 32 / (int64_t)es : 1;
  const int64_t dlda = round_up(std::max<int64_t>(ra, 1), q), dldc = round_up(n, q);
  if (k > 0) { const int e = st.reserve(0, (size_t)dlda * (size_t)std::max<int64_t>(ca, 1) * es); if (e) return e; }
  { const int e = st.reserve(2, (size_t)dldc * (size_t)n * es); if (e) return e; }
  char* dA = (char*)st.dbuf[0];
  char* dC = (char*)st.dbuf[2];
  t_h2d = t_d2h = 0;
  if (k > 0) {
    const bool paged = (size_t)ra * ca * es >= ((size_t)4 << 20) && is_pageable(hp.A);
    if (paged) { const int e = st.ring_in.h2d(dA, (size_t)dlda * es, (const char*)hp.A, (size_t)hp.lda * es, (size_t)ra * es, (size_t)ca, st.s_in); if (e) return e; }
    else B200_CUDA_TRY(cudaMemcpy2DAsync(dA, (size_t)dlda * es, hp.A, (size_t)hp.lda * es, (size_t)ra * es, (size_t)ca, cudaMemcpyHostToDevice, st.s_in));
    t_h2d += (uint64_t)ra * ca * es;
  }
  if (!beta_zero) {   // the whole window is uploaded (reading the other triangle is harmless); beta == 0: C is never read
    const bool paged = (size_t)n * n * es >= ((size_t)4 << 20) && is_pageable(hp.C);
    if (paged) { const int e = st.ring_in.h2d(dC, (size_t)dldc * es, (const char*)hp.C, (size_t)hp.ldc * es, (size_t)n * es, (size_t)n, st.s_in); if (e) return e; }
    else B200_CUDA_TRY(cudaMemcpy2DAsync(dC, (size_t)dldc * es, hp.C, (size_t)hp.ldc * es, (size_t)n * es, (size_t)n, cudaMemcpyHostToDevice, st.s_in));
    t_h2d += (uint64_t)n * n * es;
  }
  B200_CUDA_TRY(cudaEventRecord(st.ev_in[0], st.s_in));
  B200_CUDA_TRY(cudaStreamWaitEvent(st.s_comp, st.ev_in[0], 0));
  GemmProblem p = hp;
  p.A = dA; p.lda = dlda; p.B = dA; p.ldb = dlda; p.C = dC; p.ldc = dldc;
  { const int e = run_device(p, st.s_comp, B200BLAS_AUTO); if (e) { cudaDeviceSynchronize(); return e; } }
  B200_CUDA_TRY(cudaStreamSynchronize(st.s_comp));
  { const int e = ring_d2h_triangle(st.ring_out, (char*)hp.C, (size_t)hp.ldc * es, dC, (size_t)dldc * es, (size_t)n, es, hp.uplo, st.s_out); if (e) return e; }
  t_d2h += (uint64_t)n * (n + 1) / 2 * es;
  return 0;
}

static int rankk_entry(int type, bool herk, const char* uplo, const char* op, const int* pn, const int* pk, const void* palpha,
                       const void* a, const int* plda, const void* pbeta, void* c, const int* pldc) {
  NvtxRange nvtx_range("b200blas ?syrk_/?herk_");
  const bool cplx = (type == TY_C || type == TY_Z);
  const char* name = herk ? k_herk_names[type] : k_syrk_names[type];
  const int ul = (*uplo == 'U' || *uplo == 'u') ? UPLO_UPPER : (*uplo == 'L' || *uplo == 'l') ? UPLO_LOWER : -1;
  const int o = op_of(*op);
  int info = 0;
  if (ul < 0) info = 1;
  else if (o == OP_INVALID || (!herk && cplx && o == OP_C) || (herk && o == OP_T)) info = 2;
  else if (*pn < 0) info = 3;
  else if (*pk < 0) info = 4;
  else if (*plda < std::max(1, o == OP_N ? *pn : *pk)) info = 7;
  else if (*pldc < std::max(1, *pn)) info = 10;
  if (info) return xerbla_(name, &info, 6);
  GemmProblem p;
  p.type = type; p.m = *pn; p.n = *pn; p.k = *pk;
  if (herk) {   // alpha and beta are REAL scalars (blas/level3_impl.h:585-586)
    p.alpha[0] = (type == TY_C) ? (double)*(const float*)palpha : *(const double*)palpha; p.alpha[1] = 0.0;
    p.beta[0] = (type == TY_C) ? (double)*(const float*)pbeta : *(const double*)pbeta; p.beta[1] = 0.0;
  } else {
    load_scalar(type, palpha, p.alpha);
    load_scalar(type, pbeta, p.beta);
  }
  // op(A)*op(A)^T: 'N' -> A * A^T|^H, otherwise A^T|^H * A
  const int tr = herk ? OP_C : OP_T;
  if (o == OP_N) { p.opa = OP_N; p.opb = tr; } else { p.opa = tr; p.opb = OP_N; }
  p.A = a; p.lda = *plda; p.B = a; p.ldb = *plda; p.C = c; p.ldc = *pldc;
  p.uplo = ul; p.herm = herk ? 1 : 0;
  if (*pn == 0) return 0;
  const bool beta_one = (p.beta[0] == 1.0 && p.beta[1] == 0.0);
  // syrk: no product when k == 0 (:397-398); herk: also when alpha == 0 (:617).  Then only the beta pass remains, and
  // beta == 1 leaves C (including the imaginary part of a Hermitian diagonal) untouched.
  const bool product = *pk > 0 && !(herk && p.alpha[0] == 0.0);
  if (!product) {
    if (beta_one) return 0;
    p.k = 0;
  }
  t_error[0] = 0;
  int err;
  const bool dev_c = is_device_ptr(c);
  if (product && is_device_ptr(a) != dev_c) {
    snprintf(t_error, sizeof t_error, "operands must be all host or all device pointers");
    info = -1;
    return xerbla_(name, &info, 6);
  }
  if (dev_c) {
    err = run_device(p, nullptr, B200BLAS_AUTO);
    if (!err) err = fail((int)cudaStreamSynchronize(nullptr));
  } else {
    err = fail(run_host_rankk(p));
  }
  if (err) { info = -1; return xerbla_(name, &info, 6); }
  return 0;
}


// ---- the remaining level-3 routines: whole-operand staging ------------------------------------------
// ?trsm_ ?trmm_ ?symm_ ?hemm_ ?syr2k_ ?her2k_ are composites of the GEMM kernels (tri.cu, masked products); their host
// path uploads the operands whole, computes on s_comp and brings the result window back (no slab pipeline yet).
static int side_of(char x) { return (x == 'L' || x == 'l') ? 1 : (x == 'R' || x == 'r') ? 0 : -1; }
static int uplo_of(char x) { return (x == 'U' || x == 'u') ? UPLO_UPPER : (x == 'L' || x == 'l') ? UPLO_LOWER : -1; }
static int diag_of(char x) { return (x == 'U' || x == 'u') ? 1 : (x == 'N' || x == 'n') ? 0 : -1; }

// upload a rows x cols host window into staging slot `slot`; *dld receives the device leading dimension
static int stage_in(Staging& st, int slot, const void* h, int64_t ld, int64_t rows, int64_t cols, size_t es, bool copy, int64_t* dld) {
  const int64_t q = 32 / (int64_t)es > 0 ? 32 / (int64_t)es : 1;
  *dld = round_up(std::max<int64_t>(rows, 1), q);
  { const int e = st.reserve(slot, (size_t)*dld * (size_t)std::max<int64_t>(cols, 1) * es); if (e) return e; }
  if (!copy || rows == 0 || cols == 0) return 0;
  char* d = (char*)st.dbuf[slot];
  const bool paged = (size_t)rows * cols * es >= ((size_t)4 << 20) && is_pageable(h);
  if (paged) { const int e = st.ring_in.h2d(d, (size_t)*dld * es, (const char*)h, (size_t)ld * es, (size_t)rows * es, (size_t)cols, st.s_in); if (e) return e; }
  else B200_CUDA_TRY(cudaMemcpy2DAsync(d, (size_t)*dld * es, h, (size_t)ld * es, (size_t)rows * es, (size_t)cols, cudaMemcpyHostToDevice, st.s_in));
  t_h2d += (uint64_t)rows * cols * es;
  return 0;
}
static int stage_out(Staging& st, int slot, void* h, int64_t ld, int64_t rows, int64_t cols, size_t es, int64_t dld) {
  const char* d = (const char*)st.dbuf[slot];
  const bool paged = (size_t)rows * cols * es >= ((size_t)4 << 20) && is_pageable(h);
  if (paged) { const int e = st.ring_out.d2h((char*)h, (size_t)ld * es, d, (size_t)dld * es, (size_t)rows * es, (size_t)cols, st.s_out); if (e) return e; }
  else {
    B200_CUDA_TRY(cudaMemcpy2DAsync(h, (size_t)ld * es, d, (size_t)dld * es, (size_t)rows * es, (size_t)cols, cudaMemcpyDeviceToHost, st.s_out));
    B200_CUDA_TRY(cudaStreamSynchronize(st.s_out));
  }
  t_d2h += (uint64_t)rows * cols * es;
  return 0;
}
static int inputs_ready(Staging& st) {
  B200_CUDA_TRY(cudaEventRecord(st.ev_in[0], st.s_in));
  return (int)cudaStreamWaitEvent(st.s_comp, st.ev_in[0], 0);
}

static int symm_on_stream(const SymmProblem& p, cudaStream_t s) {
  const size_t ws = symm_workspace_bytes(p);
  void* w = nullptr;
  keep_pool_memory();
  { const int e = (int)cudaMallocAsync(&w, ws ? ws : 16, s); if (e) return e; }
  const int e = launch_symm(p, s, w);
  cudaFreeAsync(w, s);
  return e;
}

static const char* k_trsm_names[4] = {"STRSM ", "DTRSM ", "CTRSM ", "ZTRSM "};
static const char* k_trmm_names[4] = {"STRMM ", "DTRMM ", "CTRMM ", "ZTRMM "};

// blas/level3_impl.h:78-178 (trsm) and :183-284 (trmm)
static int tri_entry(int type, bool solve, const char* side, const char* uplo, const char* opa, const char* diag, const int* pm,
                     const int* pn, const void* palpha, const void* a, const int* plda, void* b, const int* pldb) {
  NvtxRange nvtx_range("b200blas ?trsm_/?trmm_");
  const char* name = solve ? k_trsm_names[type] : k_trmm_names[type];
  const int sd = side_of(*side), ul = uplo_of(*uplo), o = op_of(*opa), dg = diag_of(*diag);
  int info = 0;
  if (sd < 0) info = 1;
  else if (ul < 0) info = 2;
  else if (o == OP_INVALID) info = 3;
  else if (dg < 0) info = 4;
  else if (*pm < 0) info = 5;
  else if (*pn < 0) info = 6;
  else if (*plda < std::max(1, sd ? *pm : *pn)) info = 9;
  else if (*pldb < std::max(1, *pm)) info = 11;
  if (info) return xerbla_(name, &info, 6);
  const int ret = solve ? 0 : 1;   // the reference's ?trmm_ returns 1 (level3_impl.h:265,283); Fortran callers ignore it
  if (*pm == 0 || *pn == 0) return ret;
  TriProblem p;
  p.type = type; p.left = sd; p.uplo = ul; p.op = o; p.unit = dg; p.m = *pm; p.n = *pn;
  load_scalar(type, palpha, p.alpha);
  p.A = a; p.lda = *plda; p.B = b; p.ldb = *pldb;
  const bool alpha_zero = p.alpha[0] == 0.0 && p.alpha[1] == 0.0;
  t_error[0] = 0;
  int err = 0;
  const bool dev_b = is_device_ptr(b);
  if (!alpha_zero && is_device_ptr(a) != dev_b) {
    snprintf(t_error, sizeof t_error, "operands must be all host or all device pointers");
    info = -1;
    return xerbla_(name, &info, 6);
  }
  if (dev_b) {
    err = fail(solve ? launch_trsm(p, nullptr) : launch_trmm(p, nullptr));
    if (!err) err = fail((int)cudaStreamSynchronize(nullptr));
  } else {
    StageLease lease;
    if (!lease.ok()) { info = -1; return xerbla_(name, &info, 6); }
    Staging& st = lease.st();
    err = st.init();
    const size_t es = (size_t)type_bytes(type);
    const int64_t na = sd ? p.m : p.n;
    int64_t dlda = 0, dldb = 0;
    t_h2d = t_d2h = 0;
    if (!err) err = stage_in(st, 0, a, *plda, na, na, es, !alpha_zero, &dlda);
    if (!err) err = stage_in(st, 2, b, *pldb, p.m, p.n, es, !alpha_zero, &dldb);
    if (!err) err = inputs_ready(st);
    if (!err) {
      TriProblem d = p;
      d.A = st.dbuf[0]; d.lda = dlda; d.B = st.dbuf[2]; d.ldb = dldb;
      err = solve ? launch_trsm(d, st.s_comp) : launch_trmm(d, st.s_comp);
    }
    if (!err) err = (int)cudaStreamSynchronize(st.s_comp);
    if (!err) err = stage_out(st, 2, b, *pldb, p.m, p.n, es, dldb);
    if (err) cudaDeviceSynchronize();
    err = fail(err);
  }
  if (err) { info = -1; return xerbla_(name, &info, 6); }
  return ret;
}

static const char* k_symm_names[4] = {"SSYMM ", "DSYMM ", "CSYMM ", "ZSYMM "};
static const char* k_hemm_names[4] = {"", "", "CHEMM ", "ZHEMM "};

// blas/level3_impl.h:287-355 (symm) and :505-562 (hemm)
static int symm_entry(int type, bool herm, const char* side, const char* uplo, const int* pm, const int* pn, const void* palpha,
                      const void* a, const int* plda, const void* b, const int* pldb, const void* pbeta, void* c, const int* pldc) {
  NvtxRange nvtx_range("b200blas ?symm_/?hemm_");
  const char* name = herm ? k_hemm_names[type] : k_symm_names[type];
  const int sd = side_of(*side), ul = uplo_of(*uplo);
  int info = 0;
  if (sd < 0) info = 1;
  else if (ul < 0) info = 2;
  else if (*pm < 0) info = 3;
  else if (*pn < 0) info = 4;
  else if (*plda < std::max(1, sd ? *pm : *pn)) info = 7;
  else if (*pldb < std::max(1, *pm)) info = 9;
  else if (*pldc < std::max(1, *pm)) info = 12;
  if (info) return xerbla_(name, &info, 6);
  if (*pm == 0 || *pn == 0) return 1;   // level3_impl.h:316-319, :531-534
  SymmProblem p;
  p.type = type; p.left = sd; p.uplo = ul; p.herm = herm ? 1 : 0; p.m = *pm; p.n = *pn;
  load_scalar(type, palpha, p.alpha);
  load_scalar(type, pbeta, p.beta);
  p.A = a; p.lda = *plda; p.B = b; p.ldb = *pldb; p.C = c; p.ldc = *pldc;
  const bool alpha_zero = p.alpha[0] == 0.0 && p.alpha[1] == 0.0;
  const bool beta_zero = p.beta[0] == 0.0 && p.beta[1] == 0.0;
  t_error[0] = 0;
  int err = 0;
  const bool dev_c = is_device_ptr(c);
  if (!alpha_zero && (is_device_ptr(a) != dev_c || is_device_ptr(b) != dev_c)) {
    snprintf(t_error, sizeof t_error, "operands must be all host or all device pointers");
    info = -1;
    return xerbla_(name, &info, 6);
  }
  // alpha == 0: only the beta pass remains (netlib quick path) -- a k = 0 product on the GEMM kernels
  auto beta_only = [&](void* dC, int64_t dldc, cudaStream_t s) {
    GemmProblem g;
    g.type = type; g.opa = OP_N; g.opb = OP_N; g.m = p.m; g.n = p.n; g.k = 0;
    g.alpha[0] = g.alpha[1] = 0.0; g.beta[0] = p.beta[0]; g.beta[1] = p.beta[1];
    g.A = nullptr; g.lda = 1; g.B = nullptr; g.ldb = 1; g.C = dC; g.ldc = dldc;
    return run_device(g, s, B200BLAS_AUTO);
  };
  if (dev_c) {
    err = alpha_zero ? beta_only(c, *pldc, nullptr) : fail(symm_on_stream(p, nullptr));
    if (!err) err = fail((int)cudaStreamSynchronize(nullptr));
  } else {
    StageLease lease;
    if (!lease.ok()) { info = -1; return xerbla_(name, &info, 6); }
    Staging& st = lease.st();
    err = st.init();
    const size_t es = (size_t)type_bytes(type);
    const int64_t na = sd ? p.m : p.n;
    int64_t dlda = 0, dldb = 0, dldc = 0;
    t_h2d = t_d2h = 0;
    if (!err) err = stage_in(st, 0, a, *plda, na, na, es, !alpha_zero, &dlda);
    if (!err) err = stage_in(st, 1, b, *pldb, p.m, p.n, es, !alpha_zero, &dldb);
    if (!err) err = stage_in(st, 2, c, *pldc, p.m, p.n, es, !beta_zero, &dldc);
    if (!err) err = inputs_ready(st);
    if (!err) {
      SymmProblem d = p;
      d.A = st.dbuf[0]; d.lda = dlda; d.B = st.dbuf[1]; d.ldb = dldb; d.C = st.dbuf[2]; d.ldc = dldc;
      err = alpha_zero ? beta_only(st.dbuf[2], dldc, st.s_comp) : symm_on_stream(d, st.s_comp);
    }
    if (!err) err = (int)cudaStreamSynchronize(st.s_comp);
    if (!err) err = stage_out(st, 2, c, *pldc, p.m, p.n, es, dldc);
    if (err) cudaDeviceSynchronize();
    err = fail(err);
  }
  if (err) { info = -1; return xerbla_(name, &info, 6); }
  return 0;
}

static const char* k_syr2k_names[4] = {"SSYR2K", "DSYR2K", "CSYR2K", "ZSYR2K"};
static const char* k_her2k_names[4] = {"", "", "CHER2K", "ZHER2K"};

// blas/level3_impl.h:437-503 (syr2k) and :631-700 (her2k): two masked products on the GEMM kernels,
//   C.tri = alpha * op(A) op(B)^T|^H + beta * C.tri,   then   C.tri += alpha|conj(alpha) * op(B) op(A)^T|^H
static int r2k_entry(int type, bool her, const char* uplo, const char* op, const int* pn, const int* pk, const void* palpha,
                     const void* a, const int* plda, const void* b, const int* pldb, const void* pbeta, void* c, const int* pldc) {
  NvtxRange nvtx_range("b200blas ?syr2k_/?her2k_");
  const bool cplx = (type == TY_C || type == TY_Z);
  const char* name = her ? k_her2k_names[type] : k_syr2k_names[type];
  const int ul = uplo_of(*uplo), o = op_of(*op);
  int info = 0;
  if (ul < 0) info = 1;
  else if (o == OP_INVALID || (!her && cplx && o == OP_C) || (her && o == OP_T)) info = 2;
  else if (*pn < 0) info = 3;
  else if (*pk < 0) info = 4;
  else if (*plda < std::max(1, o == OP_N ? *pn : *pk)) info = 7;
  else if (*pldb < std::max(1, o == OP_N ? *pn : *pk)) info = 9;
  else if (*pldc < std::max(1, *pn)) info = 12;
  if (info) return xerbla_(name, &info, 6);
  if (*pn == 0) return 0;
  double alpha[2], beta[2];
  load_scalar(type, palpha, alpha);
  if (her) { beta[0] = (type == TY_C) ? (double)*(const float*)pbeta : *(const double*)pbeta; beta[1] = 0.0; }   // REAL beta
  else load_scalar(type, pbeta, beta);
  const bool alpha_zero = alpha[0] == 0.0 && alpha[1] == 0.0;
  const bool beta_one = beta[0] == 1.0 && beta[1] == 0.0, beta_zero = beta[0] == 0.0 && beta[1] == 0.0;
  const bool product = *pk > 0 && !alpha_zero;
  if (!product && beta_one) return her || *pk == 0 ? 1 : 0;   // C untouched (including a Hermitian diagonal)
  const int64_t n = *pn, k = product ? *pk : 0;
  const int tr = her ? OP_C : OP_T;
  t_error[0] = 0;
  const bool dev_c = is_device_ptr(c);
  if (product && (is_device_ptr(a) != dev_c || is_device_ptr(b) != dev_c)) {
    snprintf(t_error, sizeof t_error, "operands must be all host or all device pointers");
    info = -1;
    return xerbla_(name, &info, 6);
  }
  auto run = [&](const void* dA, int64_t dlda, const void* dB, int64_t dldb, void* dC, int64_t dldc, cudaStream_t s) -> int {
    GemmProblem g;
    g.type = type; g.m = n; g.n = n; g.k = k; g.uplo = ul; g.herm = her ? 1 : 0;
    if (o == OP_N) { g.opa = OP_N; g.opb = tr; } else { g.opa = tr; g.opb = OP_N; }
    g.alpha[0] = alpha[0]; g.alpha[1] = alpha[1]; g.beta[0] = beta[0]; g.beta[1] = beta[1];
    g.A = dA; g.lda = dlda; g.B = dB; g.ldb = dldb; g.C = dC; g.ldc = dldc;
    { const int e = run_device(g, s, B200BLAS_AUTO); if (e) return e; }
    if (k == 0) return 0;
    g.A = dB; g.lda = dldb; g.B = dA; g.ldb = dlda;
    if (her) g.alpha[1] = -alpha[1];
    g.beta[0] = 1.0; g.beta[1] = 0.0;
    return run_device(g, s, B200BLAS_AUTO);
  };
  int err = 0;
  if (dev_c) {
    err = run(a, *plda, b, *pldb, c, *pldc, nullptr);
    if (!err) err = fail((int)cudaStreamSynchronize(nullptr));
  } else {
    StageLease lease;
    if (!lease.ok()) { info = -1; return xerbla_(name, &info, 6); }
    Staging& st = lease.st();
    err = st.init();
    const size_t es = (size_t)type_bytes(type);
    const int64_t ra = (o == OP_N) ? n : *pk, ca = (o == OP_N) ? *pk : n;
    int64_t dlda = 0, dldb = 0, dldc = 0;
    t_h2d = t_d2h = 0;
    if (!err) err = stage_in(st, 0, a, *plda, ra, ca, es, product, &dlda);
    if (!err) err = stage_in(st, 1, b, *pldb, ra, ca, es, product, &dldb);
    if (!err) err = stage_in(st, 2, c, *pldc, n, n, es, !beta_zero, &dldc);
    if (!err) err = inputs_ready(st);
    if (!err) err = run(st.dbuf[0], dlda, st.dbuf[1], dldb, st.dbuf[2], dldc, st.s_comp);
    if (!err) err = (int)cudaStreamSynchronize(st.s_comp);
    if (!err) { err = ring_d2h_triangle(st.ring_out, (char*)c, (size_t)*pldc * es, (const char*)st.dbuf[2], (size_t)dldc * es, (size_t)n, es, ul, st.s_out); t_d2h += (uint64_t)n * (n + 1) / 2 * es; }
    if (err) cudaDeviceSynchronize();
    err = fail(err);
  }
  if (err) { info = -1; return xerbla_(name, &info, 6); }
  return (her || *pk == 0) ? 1 : 0;   // return values of level3_impl.h:470-471,502 / :672-673,699
}


// ---- ?potrf_ / ?getrf_: device-resident blocked factorizations -------------------------------------------
// lapack/cholesky.cpp:14-38 and lapack/lu.cpp:14-42.  The matrix is uploaded once, factored in HBM by lapack.cu (no
// per-block PCIe round trips, which is what makes config C5 PCIe-bound through the BLAS seam) and downloaded once.
static const char* k_potrf_names[4] = {"SPOTRF", "DPOTRF", "CPOTRF", "ZPOTRF"};
static const char* k_getrf_names[4] = {"SGETRF", "DGETRF", "CGETRF", "ZGETRF"};

// small device scratch for info + pivots, stream-ordered
struct DevInts {
  int* p = nullptr; cudaStream_t s;
  int alloc(size_t n, cudaStream_t st) { s = st; keep_pool_memory(); return (int)cudaMallocAsync((void**)&p, n * sizeof(int), st); }
  ~DevInts() { if (p) cudaFreeAsync(p, s); }
};

static int potrf_entry(int type, const char* uplo, const int* pn, void* a, const int* plda, int* info) {
  NvtxRange nvtx_range("b200blas ?potrf_");
  const int ul = uplo_of(*uplo);
  *info = 0;
  if (ul < 0) *info = -1;
  else if (*pn < 0) *info = -2;
  else if (*plda < std::max(1, *pn)) *info = -4;
  if (*info != 0) { int e = -*info; return xerbla_(k_potrf_names[type], &e, 6); }
  if (*pn == 0) return 0;
  const int64_t n = *pn;
  const size_t es = (size_t)type_bytes(type);
  t_error[0] = 0;
  int err = 0, hinfo = INT_MAX;
  const bool dev_a = is_device_ptr(a);
  StageLease lease(false);   // host operands only
  cudaStream_t s = nullptr;
  PotrfProblem p;
  p.type = type; p.uplo = ul; p.n = n; p.A = a; p.lda = *plda;
  int64_t dlda = 0;
  if (!dev_a) {
    lease.acquire();
    err = lease.ok() ? lease.st().init() : (int)cudaErrorInitializationError;
    t_h2d = t_d2h = 0;
    if (!err) err = stage_in(lease.st(), 2, a, *plda, n, n, es, true, &dlda);
    if (!err) err = inputs_ready(lease.st());
    if (!err) { s = lease.st().s_comp; p.A = lease.st().dbuf[2]; p.lda = dlda; }
  }
  DevInts di;
  if (!err) err = di.alloc(1, s);
  if (!err) err = (int)cudaMemcpyAsync(di.p, &hinfo, sizeof(int), cudaMemcpyHostToDevice, s);
  if (!err) { p.dinfo = di.p; err = launch_potrf(p, s); }
  if (!err) err = (int)cudaMemcpyAsync(&hinfo, di.p, sizeof(int), cudaMemcpyDeviceToHost, s);
  if (!err) err = (int)cudaStreamSynchronize(s);
  if (!err && !dev_a) {
    Staging& st = lease.st();
    err = ring_d2h_triangle(st.ring_out, (char*)a, (size_t)*plda * es, (const char*)st.dbuf[2], (size_t)dlda * es, (size_t)n, es, ul, st.s_out);
    t_d2h += (uint64_t)n * (n + 1) / 2 * es;
  }
  if (err) { cudaDeviceSynchronize(); fail(err); int e = -1; *info = -1; return xerbla_(k_potrf_names[type], &e, 6); }
  *info = (hinfo == INT_MAX) ? 0 : hinfo;   // index of the first non-positive pivot, 1-based (cholesky.cpp:34-35)
  return 0;
}

static int getrf_entry(int type, const int* pm, const int* pn, void* a, const int* plda, int* ipiv, int* info) {
  NvtxRange nvtx_range("b200blas ?getrf_");
  *info = 0;
  if (*pm < 0) *info = -1;
  else if (*pn < 0) *info = -2;
  else if (*plda < std::max(1, *pm)) *info = -4;
  if (*info != 0) { int e = -*info; return xerbla_(k_getrf_names[type], &e, 6); }
  if (*pm == 0 || *pn == 0) return 0;
  const int64_t m = *pm, n = *pn, size = std::min(m, n);
  const size_t es = (size_t)type_bytes(type);
  t_error[0] = 0;
  int err = 0, hinfo = INT_MAX;
  const bool dev_a = is_device_ptr(a);
  StageLease lease(false);   // host operands only
  cudaStream_t s = nullptr;
  GetrfProblem p;
  p.type = type; p.m = m; p.n = n; p.A = a; p.lda = *plda;
  int64_t dlda = 0;
  if (!dev_a) {
    lease.acquire();
    err = lease.ok() ? lease.st().init() : (int)cudaErrorInitializationError;
    t_h2d = t_d2h = 0;
    if (!err) err = stage_in(lease.st(), 2, a, *plda, m, n, es, true, &dlda);
    if (!err) err = inputs_ready(lease.st());
    if (!err) { s = lease.st().s_comp; p.A = lease.st().dbuf[2]; p.lda = dlda; }
  }
  DevInts di;
  if (!err) err = di.alloc((size_t)size + 1, s);
  if (!err) err = (int)cudaMemcpyAsync(di.p, &hinfo, sizeof(int), cudaMemcpyHostToDevice, s);
  if (!err) { p.dinfo = di.p; p.dipiv = di.p + 1; err = launch_getrf(p, s); }
  if (!err) err = (int)cudaMemcpyAsync(&hinfo, di.p, sizeof(int), cudaMemcpyDeviceToHost, s);
  if (!err) err = (int)cudaMemcpyAsync(ipiv, di.p + 1, (size_t)size * sizeof(int), cudaMemcpyDeviceToHost, s);
  if (!err) err = (int)cudaStreamSynchronize(s);
  if (!err && !dev_a) err = stage_out(lease.st(), 2, a, *plda, m, n, es, dlda);
  if (err) { cudaDeviceSynchronize(); fail(err); int e = -1; *info = -1; return xerbla_(k_getrf_names[type], &e, 6); }
  *info = (hinfo == INT_MAX) ? 0 : hinfo;   // first exactly-zero pivot, 1-based (lu.cpp:38-39)
  return 0;
}

}  // namespace b200

using namespace b200;

extern "C" {

__attribute__((weak)) int xerbla_(const char* name, int* info, int) {
  printf("Eigen BLAS ERROR #%i: %s\n", *info, name);
  return 0;
}

int sgemm_(const char* ta, const char* tb, const int* m, const int* n, const int* k, const float* alpha, const float* a,
           const int* lda, const float* b, const int* ldb, const float* beta, float* c, const int* ldc) {
  return gemm_entry(TY_S, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc);
}
int dgemm_(const char* ta, const char* tb, const int* m, const int* n, const int* k, const double* alpha,
           const double* a, const int* lda, const double* b, const int* ldb, const double* beta, double* c,
           const int* ldc) {
  return gemm_entry(TY_D, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc);
}
int cgemm_(const char* ta, const char* tb, const int* m, const int* n, const int* k, const float* alpha, const float* a,
           const int* lda, const float* b, const int* ldb, const float* beta, float* c, const int* ldc) {
  return gemm_entry(TY_C, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc);
}
int zgemm_(const char* ta, const char* tb, const int* m, const int* n, const int* k, const double* alpha,
           const double* a, const int* lda, const double* b, const int* ldb, const double* beta, double* c,
           const int* ldc) {
  return gemm_entry(TY_Z, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc);
}

#define B200_TRI(NAME, TYPE, SOLVE, RT)                                                                                            \
  int NAME(const char* side, const char* uplo, const char* opa, const char* diag, const int* m, const int* n, const RT* alpha,  \
           const RT* a, const int* lda, RT* b, const int* ldb) {                                                                \
    return tri_entry(TYPE, SOLVE, side, uplo, opa, diag, m, n, alpha, a, lda, b, ldb);                                          \
  }
B200_TRI(strsm_, TY_S, true, float) B200_TRI(dtrsm_, TY_D, true, double) B200_TRI(ctrsm_, TY_C, true, float) B200_TRI(ztrsm_, TY_Z, true, double)
B200_TRI(strmm_, TY_S, false, float) B200_TRI(dtrmm_, TY_D, false, double) B200_TRI(ctrmm_, TY_C, false, float) B200_TRI(ztrmm_, TY_Z, false, double)
#undef B200_TRI
#define B200_SYMM(NAME, TYPE, HERM, RT)                                                                                          \
  int NAME(const char* side, const char* uplo, const int* m, const int* n, const RT* alpha, const RT* a, const int* lda,         \
           const RT* b, const int* ldb, const RT* beta, RT* c, const int* ldc) {                                                 \
    return symm_entry(TYPE, HERM, side, uplo, m, n, alpha, a, lda, b, ldb, beta, c, ldc);                                        \
  }
B200_SYMM(ssymm_, TY_S, false, float) B200_SYMM(dsymm_, TY_D, false, double) B200_SYMM(csymm_, TY_C, false, float) B200_SYMM(zsymm_, TY_Z, false, double)
B200_SYMM(chemm_, TY_C, true, float) B200_SYMM(zhemm_, TY_Z, true, double)
#undef B200_SYMM
#define B200_R2K(NAME, TYPE, HER, RT)                                                                                            \
  int NAME(const char* uplo, const char* trans, const int* n, const int* k, const RT* alpha, const RT* a, const int* lda,        \
           const RT* b, const int* ldb, const RT* beta, RT* c, const int* ldc) {                                                 \
    return r2k_entry(TYPE, HER, uplo, trans, n, k, alpha, a, lda, b, ldb, beta, c, ldc);                                         \
  }
B200_R2K(ssyr2k_, TY_S, false, float) B200_R2K(dsyr2k_, TY_D, false, double) B200_R2K(csyr2k_, TY_C, false, float) B200_R2K(zsyr2k_, TY_Z, false, double)
B200_R2K(cher2k_, TY_C, true, float) B200_R2K(zher2k_, TY_Z, true, double)
#undef B200_R2K

int spotrf_(const char* uplo, const int* n, float* a, const int* lda, int* info) { return potrf_entry(TY_S, uplo, n, a, lda, info); }
int dpotrf_(const char* uplo, const int* n, double* a, const int* lda, int* info) { return potrf_entry(TY_D, uplo, n, a, lda, info); }
int cpotrf_(const char* uplo, const int* n, float* a, const int* lda, int* info) { return potrf_entry(TY_C, uplo, n, a, lda, info); }
int zpotrf_(const char* uplo, const int* n, double* a, const int* lda, int* info) { return potrf_entry(TY_Z, uplo, n, a, lda, info); }
int sgetrf_(const int* m, const int* n, float* a, const int* lda, int* ipiv, int* info) { return getrf_entry(TY_S, m, n, a, lda, ipiv, info); }
int dgetrf_(const int* m, const int* n, double* a, const int* lda, int* ipiv, int* info) { return getrf_entry(TY_D, m, n, a, lda, ipiv, info); }
int cgetrf_(const int* m, const int* n, float* a, const int* lda, int* ipiv, int* info) { return getrf_entry(TY_C, m, n, a, lda, ipiv, info); }
int zgetrf_(const int* m, const int* n, double* a, const int* lda, int* ipiv, int* info) { return getrf_entry(TY_Z, m, n, a, lda, ipiv, info); }

int ssyrk_(const char* uplo, const char* trans, const int* n, const int* k, const float* alpha, const float* a, const int* lda,
           const float* beta, float* c, const int* ldc) { return rankk_entry(TY_S, false, uplo, trans, n, k, alpha, a, lda, beta, c, ldc); }
int dsyrk_(const char* uplo, const char* trans, const int* n, const int* k, const double* alpha, const double* a, const int* lda,
           const double* beta, double* c, const int* ldc) { return rankk_entry(TY_D, false, uplo, trans, n, k, alpha, a, lda, beta, c, ldc); }
int csyrk_(const char* uplo, const char* trans, const int* n, const int* k, const float* alpha, const float* a, const int* lda,
           const float* beta, float* c, const int* ldc) { return rankk_entry(TY_C, false, uplo, trans, n, k, alpha, a, lda, beta, c, ldc); }
int zsyrk_(const char* uplo, const char* trans, const int* n, const int* k, const double* alpha, const double* a, const int* lda,
           const double* beta, double* c, const int* ldc) { return rankk_entry(TY_Z, false, uplo, trans, n, k, alpha, a, lda, beta, c, ldc); }
int cherk_(const char* uplo, const char* trans, const int* n, const int* k, const float* alpha, const float* a, const int* lda,
           const float* beta, float* c, const int* ldc) { return rankk_entry(TY_C, true, uplo, trans, n, k, alpha, a, lda, beta, c, ldc); }
int zherk_(const char* uplo, const char* trans, const int* n, const int* k, const double* alpha, const double* a, const int* lda,
           const double* beta, double* c, const int* ldc) { return rankk_entry(TY_Z, true, uplo, trans, n, k, alpha, a, lda, beta, c, ldc); }

int b200blas_gemm_dev(int type, char transa, char transb, int m, int n, int k, const void* alpha, const void* dA,
                      int64_t lda, const void* dB, int64_t ldb, const void* beta, void* dC, int64_t ldc, void* stream,
                      int variant) {
  NvtxRange nvtx_range("b200blas_gemm_dev");
  if (type < 0 || type > 3) return -1;
  const int opa = op_of(transa), opb = op_of(transb);
  int info = check_args(opa, opb, m, n, k, lda, ldb, ldc);
  if (info) return xerbla_(k_names[type], &info, 6);
  if (m == 0 || n == 0) return 0;
  GemmProblem p;
  p.type = type; p.opa = opa; p.opb = opb; p.m = m; p.n = n; p.k = k;
  load_scalar(type, alpha, p.alpha);
  load_scalar(type, beta, p.beta);
  p.A = dA; p.lda = lda; p.B = dB; p.ldb = ldb; p.C = dC; p.ldc = ldc;
  t_error[0] = 0;
  const int err = (variant == B200BLAS_AUTO && multi_wanted(p)) ? fail(multi_gemm(p, false, (cudaStream_t)stream, nullptr, nullptr))
                                                               : run_device(p, (cudaStream_t)stream, variant);
  if (err) { info = -1; return xerbla_(k_names[type], &info, 6); }
  return 0;
}

// Tensor contraction seam (unsupported/Eigen/CXX11/src/Tensor/TensorContractionCuda.h:1320-1390): the GpuDevice
// evaluator reduces every contraction to out(m x n, column-major) = lhs(m x k) * rhs(k x n) over strided views of device
// memory and then launches its own SIMT kernels.  When both views are matrices in the BLAS sense (one unit stride each) this
// entry runs the product on the sm_100a kernels instead; otherwise it returns 1 and the caller keeps its own path.
int b200blas_contract_dev(int type, int64_t m, int64_t n, int64_t k, const void* lhs, int64_t lhs_row_stride, int64_t lhs_col_stride,
                          const void* rhs, int64_t rhs_row_stride, int64_t rhs_col_stride, void* out, int64_t ldo, void* stream) {
  NvtxRange nvtx_range("b200blas_contract_dev");
  if (type < 0 || type > 3 || m < 0 || n < 0 || k < 0 || !out) return -1;
  if (m == 0 || n == 0) return 0;
  if (m > INT_MAX || n > INT_MAX || k > INT_MAX || !lhs || !rhs) return 1;
  char ta, tb;
  int64_t lda, ldb;
  // lhs(i, kk) = lhs[i * row_stride + kk * col_stride]
  if (lhs_row_stride == 1 || m == 1) { ta = 'N'; lda = std::max<int64_t>(k == 1 ? m : lhs_col_stride, std::max<int64_t>(m, 1)); if (k > 1 && lhs_col_stride < m) return 1; }
  else if (lhs_col_stride == 1 || k == 1) { ta = 'T'; lda = std::max<int64_t>(lhs_row_stride, std::max<int64_t>(k, 1)); if (lhs_row_stride < k) return 1; }
  else return 1;
  // rhs(kk, j) = rhs[kk * row_stride + j * col_stride]
  if (rhs_row_stride == 1 || k == 1) { tb = 'N'; ldb = std::max<int64_t>(n == 1 ? k : rhs_col_stride, std::max<int64_t>(k, 1)); if (n > 1 && rhs_col_stride < k) return 1; }
  else if (rhs_col_stride == 1 || n == 1) { tb = 'T'; ldb = std::max<int64_t>(rhs_row_stride, std::max<int64_t>(n, 1)); if (rhs_row_stride < n) return 1; }
  else return 1;
  const double one[2] = {1.0, 0.0}, zero[2] = {0.0, 0.0};
  float onef[2] = {1.f, 0.f}, zerof[2] = {0.f, 0.f};
  const bool sp = type == TY_S || type == TY_C;
  const int r = b200blas_gemm_dev(type, ta, tb, (int)m, (int)n, (int)k, sp ? (const void*)onef : (const void*)one, lhs, lda, rhs, ldb,
                                  sp ? (const void*)zerof : (const void*)zero, out, ldo, stream, B200BLAS_AUTO);
  if (log_enabled())
    fprintf(stderr, "[b200blas] contract %c%c m=%lld n=%lld k=%lld operands=device variant=%s%s\n", ta, tb, (long long)m, (long long)n,
            (long long)k, t_variant, r ? " FAILED" : "");
  return r == 0 ? 0 : -1;
}

int b200blas_version(void) { return 100; }
int b200blas_device_ok(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 0; }
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) { cudaGetLastError(); return 0; }
  return major == 10;
}
const char* b200blas_last_error(void) { return t_error; }
const char* b200blas_last_variant(void) { return t_variant; }
uint64_t b200blas_kernel_launches(void) { return g_launches.load(); }
void b200blas_set_variant(int variant) { g_forced_variant.store(variant == B200BLAS_AUTO ? -1 : variant); }
void b200blas_last_transfer(uint64_t* h2d, uint64_t* d2h) { if (h2d) *h2d = t_h2d; if (d2h) *d2h = t_d2h; }
void b200blas_release(void) { staging_pool().free_idle(); multi_release(); }
double b200blas_pipe_peak(int pipe, int millis) { return pipe_peak(pipe, millis); }

}  // extern "C"
