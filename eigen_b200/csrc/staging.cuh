// eigen_b200/csrc/staging.cuh -- host-side staging machinery shared by host.cu (single-device calls) and multi.cu (the
// multi-device partitioner): the copy-worker pool, the pinned ring for pageable operands, and the per-call staging
// contexts (streams, events, grow-only device images).  Contexts are POOLED per device: concurrent callers (the
// reference's ?gemm_ is re-entrant) each lease their own context instead of serialising on one mutex.
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"

namespace b200 {

// ---- pageable operands: pinned staging ring + copy workers -------------------------------------------------------
// Eigen matrices are ordinary pageable memory.  A cudaMemcpy2DAsync from pageable memory is staged by the driver in
// small synchronous pieces (~11 GB/s measured through bench_gemm -DHAVE_BLAS).  Instead, worker threads copy the
// operand into a ring of pinned buffers (several threads are needed to outrun one PCIe Gen5 x16 link) and each filled
// buffer goes to the device with one asynchronous 2-D DMA while the workers fill the next one.
class CopyPool {
 public:
  explicit CopyPool(int nthreads) : stop_(false), gen_(0), pending_(0) {
    for (int i = 0; i < nthreads; ++i) th_.emplace_back([this] { worker(); });
  }
  ~CopyPool() {
    { std::lock_guard<std::mutex> l(mu_); stop_ = true; ++gen_; }
    cv_.notify_all();
    for (auto& t : th_) t.join();
  }
  // dst/src 2-D regions of `ncols` columns of `width` bytes with the given pitches; returns when the copy is done
  void copy2d(char* dst, size_t dpitch, const char* src, size_t spitch, size_t width, size_t ncols) {
    if (width * ncols < (1u << 20) || th_.empty()) {
      for (size_t j = 0; j < ncols; ++j) memcpy(dst + j * dpitch, src + j * spitch, width);
      return;
    }
    std::unique_lock<std::mutex> call(call_mu_);  // one parallel copy at a time
    {
      std::lock_guard<std::mutex> l(mu_);
      job_ = {dst, dpitch, src, spitch, width, ncols};
      // split every column into pieces of <= 1 MiB so that tall-and-thin regions parallelise as well
      piece_ = width > (1u << 20) ? (1u << 20) : width;
      pieces_per_col_ = (width + piece_ - 1) / piece_;
      next_.store(0);
      total_ = ncols * pieces_per_col_;
      pending_ = (int)th_.size();
      ++gen_;
    }
    cv_.notify_all();
    run();  // the caller works too
    std::unique_lock<std::mutex> l(mu_);
    done_cv_.wait(l, [this] { return pending_ == 0; });
  }

 private:
  struct Job { char* dst; size_t dpitch; const char* src; size_t spitch; size_t width; size_t ncols; };
  void run() {
    for (;;) {
      const size_t i = next_.fetch_add(1);
      if (i >= total_) break;
      const size_t col = i / pieces_per_col_, off = (i % pieces_per_col_) * piece_;
      const size_t len = off + piece_ <= job_.width ? piece_ : job_.width - off;
      memcpy(job_.dst + col * job_.dpitch + off, job_.src + col * job_.spitch + off, len);
    }
  }
  void worker() {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> l(mu_);
        cv_.wait(l, [&] { return gen_ != seen; });
        seen = gen_;
        if (stop_) return;
      }
      run();
      {
        std::lock_guard<std::mutex> l(mu_);
        if (--pending_ == 0) done_cv_.notify_all();
      }
    }
  }
  std::vector<std::thread> th_;
  std::mutex mu_, call_mu_;
  std::condition_variable cv_, done_cv_;
  bool stop_;
  uint64_t gen_;
  int pending_;
  Job job_{};
  size_t piece_ = 0, pieces_per_col_ = 1, total_ = 0;
  std::atomic<size_t> next_{0};
};

inline CopyPool& copy_pool() {
  static CopyPool pool([] {
    const char* e = getenv("B200BLAS_COPY_THREADS");
    int n = e ? atoi(e) : 0;
    if (n <= 0) {
      const unsigned hc = std::thread::hardware_concurrency();
      // the caller is blocked inside a BLAS call, so its cores are free to copy: measured on a 16-core B200 host
      // (profiles/bench_r02/e2e_copy_threads_pass5.jsonl, dgemm 16384^3 from pageable memory): 7 threads 27.1, 12 threads 29.5,
      // 16 threads 29.9 TFLOP/s
      n = hc >= 16 ? 16 : hc >= 8 ? (int)hc - 2 : hc >= 4 ? 3 : 1;
    }
    return n - 1 < 0 ? 0 : n - 1;  // the calling thread is the n-th copier
  }());
  return pool;
}

inline bool is_pageable(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return true; }
  return at.type == cudaMemoryTypeUnregistered;
}

struct PinnedRing;
inline int ring_d2h_triangle(PinnedRing& ring, char* dst, size_t dpitch, const char* src, size_t spitch, size_t n, size_t es,
                             int uplo, cudaStream_t s);
struct PinnedRing {
  static constexpr int NBUF = 4;
  static constexpr size_t BYTES = (size_t)32 << 20;
  static constexpr size_t SMALL_BYTES = (size_t)4 << 20;
  char* buf[NBUF] = {nullptr, nullptr, nullptr, nullptr};
  char* small = nullptr;           // bounce buffer for small triangular results
  cudaEvent_t free_ev[NBUF] = {nullptr, nullptr, nullptr, nullptr};
  bool used[NBUF] = {false, false, false, false};
  int next = 0;
  bool ready = false;
  int init() {
    if (ready) return 0;
    for (int i = 0; i < NBUF; ++i) {
      if (!buf[i]) B200_CUDA_TRY(cudaHostAlloc((void**)&buf[i], BYTES, cudaHostAllocDefault));
      if (!free_ev[i]) B200_CUDA_TRY(cudaEventCreateWithFlags(&free_ev[i], cudaEventDisableTiming));
    }
    ready = true;
    return 0;
  }
  int init_small() {
    if (small) return 0;
    return (int)cudaHostAlloc((void**)&small, SMALL_BYTES, cudaHostAllocDefault);
  }
  void release() {   // also tears down a partially initialised ring
    for (int i = 0; i < NBUF; ++i) {
      if (buf[i]) cudaFreeHost(buf[i]);
      if (free_ev[i]) cudaEventDestroy(free_ev[i]);
      buf[i] = nullptr; free_ev[i] = nullptr; used[i] = false;
    }
    if (small) { cudaFreeHost(small); small = nullptr; }
    ready = false;
  }
  // host (pageable) -> device, 2-D, through the ring; asynchronous with respect to the device stream
  int h2d(char* dst, size_t dpitch, const char* src, size_t spitch, size_t width, size_t ncols, cudaStream_t s) {
    if (width == 0 || ncols == 0) return 0;
    if (width > BYTES) return (int)cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, ncols, cudaMemcpyHostToDevice, s);
    { const int e = init(); if (e) return e; }
    const size_t cols_per = BYTES / width;
    for (size_t c0 = 0; c0 < ncols; c0 += cols_per) {
      const size_t nc = ncols - c0 < cols_per ? ncols - c0 : cols_per;
      const int slot = next;
      next = (next + 1) % NBUF;
      if (used[slot]) B200_CUDA_TRY(cudaEventSynchronize(free_ev[slot]));
      copy_pool().copy2d(buf[slot], width, src + c0 * spitch, spitch, width, nc);
      B200_CUDA_TRY(cudaMemcpy2DAsync(dst + c0 * dpitch, dpitch, buf[slot], width, width, nc, cudaMemcpyHostToDevice, s));
      B200_CUDA_TRY(cudaEventRecord(free_ev[slot], s));
      used[slot] = true;
    }
    return 0;
  }
  // device -> host (pageable), 2-D, through the ring; returns when the data is in the caller's memory
  int d2h(char* dst, size_t dpitch, const char* src, size_t spitch, size_t width, size_t ncols, cudaStream_t s) {
    if (width == 0 || ncols == 0) return 0;
    if (width > BYTES) {
      B200_CUDA_TRY(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, ncols, cudaMemcpyDeviceToHost, s));
      return (int)cudaStreamSynchronize(s);
    }
    { const int e = init(); if (e) return e; }
    const size_t cols_per = BYTES / width;
    // software pipeline over the ring: DMA of chunk i+1 runs while the workers copy chunk i out
    size_t issued = 0, retired = 0;
    const size_t nchunks = (ncols + cols_per - 1) / cols_per;
    while (retired < nchunks) {
      while (issued < nchunks && issued - retired < (size_t)NBUF) {
        const size_t c0 = issued * cols_per, nc = ncols - c0 < cols_per ? ncols - c0 : cols_per;
        const int slot = (int)(issued % NBUF);
        B200_CUDA_TRY(cudaMemcpy2DAsync(buf[slot], width, src + c0 * spitch, spitch, width, nc, cudaMemcpyDeviceToHost, s));
        B200_CUDA_TRY(cudaEventRecord(free_ev[slot], s));
        used[slot] = false;
        ++issued;
      }
      const size_t c0 = retired * cols_per, nc = ncols - c0 < cols_per ? ncols - c0 : cols_per;
      const int slot = (int)(retired % NBUF);
      B200_CUDA_TRY(cudaEventSynchronize(free_ev[slot]));
      copy_pool().copy2d(dst + c0 * dpitch, dpitch, buf[slot], width, width, nc);
      ++retired;
    }
    return 0;
  }
};

// device -> host for a rank-k update: whole columns travel into the pinned ring, but only the referenced triangle of
// each column is copied into the caller's matrix (the other triangle is not referenced by ?syrk_/?herk_ and may be
// in use by the caller).  Returns when the data is in place.
inline int ring_d2h_triangle(PinnedRing& ring, char* dst, size_t dpitch, const char* src, size_t spitch, size_t n, size_t es,
                             int uplo, cudaStream_t s) {
  if (n == 0) return 0;
  const size_t width = n * es;
  if (width * n <= PinnedRing::SMALL_BYTES) {
    // small result: one 2-D DMA into a 4 MiB pinned bounce buffer (the 4 x 32 MiB ring is only set up for large outputs)
    { const int e = ring.init_small(); if (e) return e; }
    B200_CUDA_TRY(cudaMemcpy2DAsync(ring.small, width, src, spitch, width, n, cudaMemcpyDeviceToHost, s));
    B200_CUDA_TRY(cudaStreamSynchronize(s));
    for (size_t j = 0; j < n; ++j) {
      const size_t lo = uplo == UPLO_UPPER ? 0 : j, hi = uplo == UPLO_UPPER ? j + 1 : n;
      memcpy(dst + j * dpitch + lo * es, ring.small + j * width + lo * es, (hi - lo) * es);
    }
    return 0;
  }
  { const int e = ring.init(); if (e) return e; }
  const size_t cols_per = width > PinnedRing::BYTES ? 0 : PinnedRing::BYTES / width;
  if (cols_per == 0) {
    // a single column exceeds a ring buffer (n > 4M doubles): copy the triangle column by column
    for (size_t j = 0; j < n; ++j) {
      const size_t lo = uplo == UPLO_UPPER ? 0 : j, hi = uplo == UPLO_UPPER ? j + 1 : n;
      B200_CUDA_TRY(cudaMemcpyAsync(dst + j * dpitch + lo * es, src + j * spitch + lo * es, (hi - lo) * es, cudaMemcpyDeviceToHost, s));
    }
    return (int)cudaStreamSynchronize(s);
  }
  const size_t nchunks = (n + cols_per - 1) / cols_per;
  size_t issued = 0, retired = 0;
  while (retired < nchunks) {
    while (issued < nchunks && issued - retired < (size_t)PinnedRing::NBUF) {
      const size_t c0 = issued * cols_per, nc = n - c0 < cols_per ? n - c0 : cols_per;
      const int slot = (int)(issued % PinnedRing::NBUF);
      B200_CUDA_TRY(cudaMemcpy2DAsync(ring.buf[slot], width, src + c0 * spitch, spitch, width, nc, cudaMemcpyDeviceToHost, s));
      B200_CUDA_TRY(cudaEventRecord(ring.free_ev[slot], s));
      ++issued;
    }
    const size_t c0 = retired * cols_per, nc = n - c0 < cols_per ? n - c0 : cols_per;
    const int slot = (int)(retired % PinnedRing::NBUF);
    B200_CUDA_TRY(cudaEventSynchronize(ring.free_ev[slot]));
    for (size_t jj = 0; jj < nc; ++jj) {
      const size_t j = c0 + jj;
      const size_t lo = uplo == UPLO_UPPER ? 0 : j, hi = uplo == UPLO_UPPER ? j + 1 : n;
      memcpy(dst + j * dpitch + lo * es, ring.buf[slot] + jj * width + lo * es, (hi - lo) * es);
    }
    ++retired;
  }
  return 0;
}

// ---- staging context: streams, events and grow-only device images of one in-flight host-operand call -----------
struct Staging {
  bool leased = false;   // owned by StagingPool
  int dev = -1;
  cudaStream_t s_in = nullptr, s_comp = nullptr, s_out = nullptr;
  void* dbuf[3] = {nullptr, nullptr, nullptr};
  size_t dcap[3] = {0, 0, 0};
  static constexpr int MAX_SLABS = 64;
  static constexpr int MAX_ACHUNKS = 16;
  cudaEvent_t ev_in[MAX_SLABS] = {}, ev_comp[MAX_SLABS] = {}, ev_ac[MAX_ACHUNKS] = {}, ev_a = nullptr;
  bool ready = false;
  PinnedRing ring_in, ring_out;   // pageable operands only

  int init() {
    int d = 0;
    B200_CUDA_TRY(cudaGetDevice(&d));
    if (ready && d == dev) return 0;
    release();
    dev = d;
    B200_CUDA_TRY(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
    B200_CUDA_TRY(cudaStreamCreateWithFlags(&s_comp, cudaStreamNonBlocking));
    B200_CUDA_TRY(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
    B200_CUDA_TRY(cudaEventCreateWithFlags(&ev_a, cudaEventDisableTiming));
    for (int i = 0; i < MAX_SLABS; ++i) {
      B200_CUDA_TRY(cudaEventCreateWithFlags(&ev_in[i], cudaEventDisableTiming));
      B200_CUDA_TRY(cudaEventCreateWithFlags(&ev_comp[i], cudaEventDisableTiming));
    }
    for (int i = 0; i < MAX_ACHUNKS; ++i) B200_CUDA_TRY(cudaEventCreateWithFlags(&ev_ac[i], cudaEventDisableTiming));
    ready = true;
    return 0;
  }
  int reserve(int i, size_t bytes) {
    if (bytes <= dcap[i]) return 0;
    if (dbuf[i]) { cudaFree(dbuf[i]); dbuf[i] = nullptr; dcap[i] = 0; }
    B200_CUDA_TRY(cudaMalloc(&dbuf[i], bytes));
    dcap[i] = bytes;
    return 0;
  }
  void release() {   // also tears down a partially initialised context (init() failed half-way)
    for (int i = 0; i < 3; ++i) { if (dbuf[i]) cudaFree(dbuf[i]); dbuf[i] = nullptr; dcap[i] = 0; }
    if (s_in) cudaStreamDestroy(s_in);
    if (s_comp) cudaStreamDestroy(s_comp);
    if (s_out) cudaStreamDestroy(s_out);
    if (ev_a) cudaEventDestroy(ev_a);
    for (int i = 0; i < MAX_SLABS; ++i) {
      if (ev_in[i]) cudaEventDestroy(ev_in[i]);
      if (ev_comp[i]) cudaEventDestroy(ev_comp[i]);
      ev_in[i] = ev_comp[i] = nullptr;
    }
    for (int i = 0; i < MAX_ACHUNKS; ++i) { if (ev_ac[i]) cudaEventDestroy(ev_ac[i]); ev_ac[i] = nullptr; }
    ring_in.release();
    ring_out.release();
    s_in = s_comp = s_out = nullptr; ev_a = nullptr;
    ready = false;
  }
};

// ---- pooled staging contexts ------------------------------------------------------------------------------------------
// acquire() hands out the most recently released idle context of the current device (so a single-threaded caller keeps
// hitting the same warm buffers) and creates a new one when all are busy, up to MAX_PER_DEVICE; beyond that callers wait.
class StagingPool {
 public:
  static constexpr int MAX_PER_DEVICE = 4;
  Staging* acquire() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    std::unique_lock<std::mutex> l(mu_);
    for (;;) {
      int busy = 0;
      for (auto it = ctx_.rbegin(); it != ctx_.rend(); ++it) {
        if ((*it)->dev != dev && (*it)->dev != -1) continue;
        if (!(*it)->leased) { (*it)->leased = true; Staging* st = it->get(); touch(st); return st; }
        ++busy;
      }
      if (busy < MAX_PER_DEVICE) {
        ctx_.emplace_back(new Staging());
        ctx_.back()->leased = true;
        return ctx_.back().get();
      }
      cv_.wait(l);
    }
  }
  void release(Staging* st) {
    { std::lock_guard<std::mutex> l(mu_); st->leased = false; }
    cv_.notify_one();
  }
  void free_idle() {   // b200blas_release: drop the cached device images / pinned rings of every idle context
    std::lock_guard<std::mutex> l(mu_);
    for (auto& c : ctx_) if (!c->leased) { int cur = 0; cudaGetDevice(&cur); if (c->dev >= 0) cudaSetDevice(c->dev); c->release(); cudaSetDevice(cur); }
  }
 private:
  void touch(Staging* st) {   // move to the back: most recently used first in acquire()
    for (size_t i = 0; i < ctx_.size(); ++i)
      if (ctx_[i].get() == st) { auto p = std::move(ctx_[i]); ctx_.erase(ctx_.begin() + i); ctx_.push_back(std::move(p)); return; }
  }
  std::mutex mu_;
  std::condition_variable cv_;
  std::vector<std::unique_ptr<Staging>> ctx_;
};
inline StagingPool& staging_pool() { static StagingPool p; return p; }

// RAII lease of one staging context of the current device
class StageLease {
 public:
  explicit StageLease(bool now = true) { if (now) acquire(); }
  ~StageLease() { if (st_) staging_pool().release(st_); }
  StageLease(const StageLease&) = delete;
  StageLease& operator=(const StageLease&) = delete;
  void acquire() { if (!st_) st_ = staging_pool().acquire(); }
  bool ok() const { return st_ != nullptr; }
  Staging& st() { return *st_; }
 private:
  Staging* st_ = nullptr;
};

inline bool is_device_ptr(const void* p) {
  if (!p) return false;
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

inline int64_t round_up(int64_t x, int64_t q) { return (x + q - 1) / q * q; }

}  // namespace b200
