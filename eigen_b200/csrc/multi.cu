// eigen_b200/csrc/multi.cu -- the multi-GPU partitioner behind ?gemm_ / b200blas_gemm_dev (include/b200blas.h section 3).
//
// B200 counterpart of parallelize_gemm + GemmParallelInfo (Eigen/src/Core/products/Parallelizer.h:74-157) and of the
// threaded branch of general_matrix_matrix_product::run (GeneralMatrixMatrix.h:83-152).  The reference cuts C into one
// column slab per OpenMP thread inside the product call; every thread reads all of A and its own columns of B and the
// threads pack A' cooperatively (each packs 1/T of every kc block, the others wait on info[].sync / info[].users).
// Here the workers are GPUs driven by ONE host process:
//   * C is cut into a pr x pc grid of tiles (k is never split, so there is no reduction);
//   * "cooperative packing" becomes cooperative FETCHING: every k-chunk of a panel A_i / B_j crosses the slow link
//     (PCIe from host memory, or the root GPU's NVLink egress) exactly once, to one owner device of its grid row /
//     column, and is relayed from there to the other devices of the row / column by peer-to-peer copies on the copy
//     engines (cudaMemcpy2DAsync between devices with peer access; no SM is involved, unlike NCCL's kernels);
//   * the product on the chunks that have landed overlaps the transfer of the later ones (groups of 1, 1, 2, 4, 8 ...
//     chunks, each a full-tile launch on the single-GPU kernels); the last group runs column sub-slab by sub-slab and
//     finished sub-slabs flow back to the caller's C while the next one computes; beta*C is folded in by a small
//     bandwidth-bound kernel (never read when beta == 0).
// The partition is built as DATA (b200blas_multi_plan: a list of copy / product / axpby steps with stream slots and
// dependencies) and then executed on CUDA streams; tests/test_multi_plan.py interprets the same plan with numpy on
// the CPU and checks both the result and that every pair of conflicting steps is ordered by a dependency.
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/b200blas.h"
#include "common.cuh"
#include "staging.cuh"

namespace b200 {
namespace {

constexpr int MAXDEV = B200BLAS_PLAN_MAXDEV;
constexpr int MAXCH = B200BLAS_PLAN_MAXCHUNK;
enum { SLOT_FETCH = 0, SLOT_RELAY = 1, SLOT_COMP = 2, SLOT_RET = 3, SLOT_FOLD = 4, NSLOT = 5 };
enum { BUF_A = 0, BUF_B = 1, BUF_P = 2, BUF_CIN = 3, BUF_RECV0 = 4 };

std::atomic<int> g_ndev{-1};          // -1: read B200BLAS_NGPUS on first use
std::atomic<int> g_grid_pr{0}, g_grid_pc{0};

int visible_devices() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}
// B200BLAS_MULTI_VIRTUAL=1 (tests): more plan devices than physical ones are allowed; plan device w runs on physical
// device (root + w) mod visible with its own streams and buffers, so the whole executor can be exercised on ONE GPU.
bool virtual_devices_ok() {
  const char* e = getenv("B200BLAS_MULTI_VIRTUAL");
  return e && e[0] && e[0] != '0';
}
int clamp_devices(int n) {
  if (n < 1) n = 1;
  if (n > 1 && !virtual_devices_ok()) { const int v = visible_devices(); if (n > v) n = v > 0 ? v : 1; }
  return n > MAXDEV ? MAXDEV : n;
}

int devices_in_effect() {
  int n = g_ndev.load();
  if (n < 0) {
    const char* e = getenv("B200BLAS_NGPUS");
    n = clamp_devices(e ? atoi(e) : 1);
    g_ndev.store(n);
  }
  return n;
}

void default_grid(int ndev, int* pr, int* pc) {
  int r = g_grid_pr.load(), c = g_grid_pc.load();
  if (r <= 0 || c <= 0 || r * c != ndev) {
    static const int env_r = [] { const char* e = getenv("B200BLAS_GRID"); int a = 0, b = 0; return (e && sscanf(e, "%dx%d", &a, &b) == 2) ? a : 0; }();
    static const int env_c = [] { const char* e = getenv("B200BLAS_GRID"); int a = 0, b = 0; return (e && sscanf(e, "%dx%d", &a, &b) == 2) ? b : 0; }();
    r = env_r; c = env_c;
  }
  if (r <= 0 || c <= 0 || r * c != ndev) {
    // 2 -> 1x2, 4 -> 2x2, 8 -> 2x4: the most square grid with pr <= pc (row cuts are strided in a
    // column-major operand, column cuts are contiguous, so the longer side goes to the columns)
    r = 1;
    for (int d = 1; d * d <= ndev; ++d) if (ndev % d == 0) r = d;
    c = ndev / r;
  }
  *pr = r; *pc = c;
}

// [0, extent) in `parts` consecutive ranges, lengths multiples of `quantum` except the last (the rule of
// Parallelizer.h:140-151 -- blockCols & ~3, blockRows rounded to mr -- with the pair-tile edge as quantum)
void split_cuts(int64_t extent, int parts, int64_t quantum, int64_t* cuts) {
  int64_t block = (extent + parts - 1) / parts;
  block = (block + quantum - 1) / quantum * quantum;
  for (int p = 0; p <= parts; ++p) cuts[p] = std::min<int64_t>(extent, (int64_t)p * block);
  cuts[parts] = extent;
}

struct Builder {
  b200blas_plan_info* info;
  b200blas_step* steps;
  int cap;
  int n = 0;
  int last_on[MAXDEV][NSLOT];
  Builder(b200blas_plan_info* i, b200blas_step* s, int c) : info(i), steps(s), cap(c) {
    for (auto& d : last_on) for (int& x : d) x = -1;
  }
  int add(const b200blas_step& st_in) {
    b200blas_step st = st_in;
    // a wait on a step of the SAME stream is implied by stream order; duplicates are dropped
    int w = 0;
    for (int i = 0; i < st_in.nwait; ++i) {
      const int idx = st_in.wait[i];
      if (idx < 0) continue;
      bool dup = false;
      for (int j = 0; j < w; ++j) dup = dup || st.wait[j] == idx;
      if (dup) continue;
      if (idx < cap && steps[idx].dev == st.dev && steps[idx].stream == st.stream) continue;
      st.wait[w++] = idx;
    }
    st.nwait = w;
    for (int i = w; i < 4; ++i) st.wait[i] = -1;
    st.record = 0;
    if (n < cap) {
      steps[n] = st;
      for (int i = 0; i < w; ++i) if (st.wait[i] < cap) steps[st.wait[i]].record = 1;
    }
    last_on[st.dev][st.stream] = n;
    return n++;
  }
};

b200blas_region region(int loc, int buf, int64_t r0, int64_t c0, int64_t rows, int64_t cols) {
  b200blas_region r;
  r.loc = loc; r.buf = buf; r.r0 = r0; r.c0 = c0; r.rows = rows; r.cols = cols;
  return r;
}
b200blas_step blank_step(int kind, int dev, int stream) {
  b200blas_step s;
  memset(&s, 0, sizeof s);
  s.kind = kind; s.dev = dev; s.stream = stream;
  s.x = s.y = s.z = region(-2, 0, 0, 0, 0, 0);
  s.alpha[0] = 1.0; s.beta[0] = 0.0;
  for (int& w : s.wait) w = -1;
  return s;
}

// column sub-slabs of a tile for the return pipeline: 1/2 + 3/8 + 1/8 so that only an eighth of the tile is still
// on the wire when the last product finishes
int subslab_cuts(int64_t width, int64_t* cuts) {
  const int64_t q = 256;
  if (width >= 16 * q) {
    cuts[0] = 0; cuts[1] = (width / 2) / q * q; cuts[2] = (width * 7 / 8) / q * q; cuts[3] = width;
    return 3;
  }
  if (width >= 4 * q) { cuts[0] = 0; cuts[1] = (width / 2) / q * q; cuts[2] = width; return 2; }
  cuts[0] = 0; cuts[1] = width;
  return width > 0 ? 1 : 0;
}

int build_plan(int type, int opa, int opb, int64_t m, int64_t n, int64_t k, const double* alpha, const double* beta,
               int ndev, int pr, int pc, bool host_origin, b200blas_plan_info* info, b200blas_step* steps, int cap) {
  if (ndev < 1 || ndev > MAXDEV || pr * pc != ndev || m <= 0 || n <= 0 || k <= 0) return -1;
  memset(info, 0, sizeof *info);
  info->ndev = ndev; info->pr = pr; info->pc = pc; info->host_origin = host_origin ? 1 : 0;
  const int64_t es = type_bytes(type);
  const int64_t q = std::max<int64_t>(1, 32 / es);   // device leading dimensions: multiples of 32 bytes
  split_cuts(m, pr, 256, info->row_cut);
  split_cuts(n, pc, 256, info->col_cut);
  // k-chunks: ~16 equal chunks, multiples of 256, at least 512 long
  int64_t kc = std::max<int64_t>(512, ((k + 15) / 16 + 255) / 256 * 256);
  int nch = (int)((k + kc - 1) / kc);
  if (nch > MAXCH) { kc = ((k + MAXCH - 1) / MAXCH + 255) / 256 * 256; nch = (int)((k + kc - 1) / kc); }
  info->nchunks = nch;
  for (int c = 0; c <= nch; ++c) info->k_cut[c] = std::min<int64_t>(k, (int64_t)c * kc);
  // groups of 1, 1, 2, 4, 8 ... chunks: the first products start after 1/16 of the traffic, later ones amortise the
  // per-launch costs (pipeline fill, C read-modify-write, wave tails) over more work
  int ng = 0;
  for (int c = 0, sz = 1, first = 1; c < nch;) {
    info->group_first_chunk[ng++] = c;
    c += sz;
    if (first) first = 0; else sz *= 2;
  }
  // a short tail group is merged into its predecessor
  if (ng >= 2 && nch - info->group_first_chunk[ng - 1] < (info->group_first_chunk[ng - 1] - info->group_first_chunk[ng - 2] + 1) / 2) --ng;
  info->group_first_chunk[ng] = nch;
  info->ngroups = ng;
  const bool beta_zero = beta[0] == 0.0 && beta[1] == 0.0;
  const int root = 0;
  auto dev_of = [&](int i, int j) { return i * pc + j; };
  auto uses_origin = [&](int d) { return !host_origin && d == root; };   // the root multiplies straight out of the caller's matrices
  for (int i = 0; i < pr; ++i)
    for (int j = 0; j < pc; ++j) {
      const int d = dev_of(i, j);
      const int64_t mi = info->row_cut[i + 1] - info->row_cut[i], nj = info->col_cut[j + 1] - info->col_cut[j];
      if (uses_origin(d) || mi <= 0 || nj <= 0) continue;
      info->ld[d][BUF_A] = opa == OP_N ? round_up(mi, q) : round_up(k, q);
      info->elems[d][BUF_A] = info->ld[d][BUF_A] * (opa == OP_N ? k : mi);
      info->ld[d][BUF_B] = opb == OP_N ? round_up(k, q) : round_up(nj, q);
      info->elems[d][BUF_B] = info->ld[d][BUF_B] * (opb == OP_N ? nj : k);
      info->ld[d][BUF_P] = round_up(mi, q);
      info->elems[d][BUF_P] = info->ld[d][BUF_P] * nj;
      if (host_origin && !beta_zero) { info->ld[d][BUF_CIN] = info->ld[d][BUF_P]; info->elems[d][BUF_CIN] = info->elems[d][BUF_P]; }
    }
  Builder B(info, steps, cap);
  // step that delivered chunk c of A_i / B_j to device d (-1: the device reads the origin / nothing was needed)
  std::vector<int> a_here((size_t)ndev * nch, -1), b_here((size_t)ndev * nch, -1);
  auto tile_empty = [&](int d) {
    const int i = d / pc, j = d % pc;
    return info->row_cut[i + 1] <= info->row_cut[i] || info->col_cut[j + 1] <= info->col_cut[j];
  };
  auto a_origin = [&](int i, int64_t k0, int64_t kk) {
    const int64_t r0 = info->row_cut[i], mi = info->row_cut[i + 1] - r0;
    return opa == OP_N ? region(-1, BUF_A, r0, k0, mi, kk) : region(-1, BUF_A, k0, r0, kk, mi);
  };
  auto a_local = [&](int d, int64_t k0, int64_t kk) {
    const int i = d / pc;
    const int64_t mi = info->row_cut[i + 1] - info->row_cut[i];
    return opa == OP_N ? region(d, BUF_A, 0, k0, mi, kk) : region(d, BUF_A, k0, 0, kk, mi);
  };
  auto b_origin = [&](int j, int64_t s0, int64_t ns, int64_t k0, int64_t kk) {
    const int64_t c0 = info->col_cut[j] + s0;
    return opb == OP_N ? region(-1, BUF_B, k0, c0, kk, ns) : region(-1, BUF_B, c0, k0, ns, kk);
  };
  auto b_local = [&](int d, int64_t s0, int64_t ns, int64_t k0, int64_t kk) {
    return opb == OP_N ? region(d, BUF_B, k0, s0, kk, ns) : region(d, BUF_B, s0, k0, ns, kk);
  };
  auto emit_chunk = [&](int c) {
    const int64_t k0 = info->k_cut[c], kk = info->k_cut[c + 1] - k0;
    // ---- A_i: owner = device (i, c mod pc) ----
    for (int i = 0; i < pr; ++i) {
      if (info->row_cut[i + 1] <= info->row_cut[i]) continue;
      int jo = c % pc;
      for (int t = 0; t < pc && tile_empty(dev_of(i, jo)); ++t) jo = (jo + 1) % pc;   // skip devices without a tile
      const int owner = dev_of(i, jo);
      if (tile_empty(owner)) continue;
      int fetch = -1;
      if (!uses_origin(owner)) {
        b200blas_step s = blank_step(B200BLAS_STEP_COPY, owner, SLOT_FETCH);
        s.x = a_origin(i, k0, kk); s.z = a_local(owner, k0, kk);
        fetch = B.add(s);
        a_here[(size_t)owner * nch + c] = fetch;
      }
      for (int j = 0; j < pc; ++j) {
        const int d = dev_of(i, j);
        if (d == owner || uses_origin(d) || tile_empty(d)) continue;
        b200blas_step s = blank_step(B200BLAS_STEP_COPY, d, SLOT_RELAY);
        s.x = uses_origin(owner) ? a_origin(i, k0, kk) : a_local(owner, k0, kk);
        s.z = a_local(d, k0, kk);
        s.nwait = 1; s.wait[0] = fetch;
        a_here[(size_t)d * nch + c] = B.add(s);
      }
    }
    // ---- B_j: owner = device (c mod pr, j) ----
    for (int j = 0; j < pc; ++j) {
      const int64_t nj = info->col_cut[j + 1] - info->col_cut[j];
      if (nj <= 0) continue;
      int io = c % pr;
      for (int t = 0; t < pr && tile_empty(dev_of(io, j)); ++t) io = (io + 1) % pr;
      const int owner = dev_of(io, j);
      if (tile_empty(owner)) continue;
      int fetch = -1;
      if (!uses_origin(owner)) {
        b200blas_step s = blank_step(B200BLAS_STEP_COPY, owner, SLOT_FETCH);
        s.x = b_origin(j, 0, nj, k0, kk); s.z = b_local(owner, 0, nj, k0, kk);
        fetch = B.add(s);
        b_here[(size_t)owner * nch + c] = fetch;
      }
      for (int i = 0; i < pr; ++i) {
        const int d = dev_of(i, j);
        if (d == owner || uses_origin(d) || tile_empty(d)) continue;
        b200blas_step s = blank_step(B200BLAS_STEP_COPY, d, SLOT_RELAY);
        s.x = uses_origin(owner) ? b_origin(j, 0, nj, k0, kk) : b_local(owner, 0, nj, k0, kk);
        s.z = b_local(d, 0, nj, k0, kk);
        s.nwait = 1; s.wait[0] = fetch;
        b_here[(size_t)d * nch + c] = B.add(s);
      }
    }
  };
  // one product of device d: tile columns [s0, s0 + ns), chunks [c0, c1)
  auto emit_gemm = [&](int d, int64_t s0, int64_t ns, int c0, int c1, bool first) {
    const int i = d / pc, j = d % pc;
    const int64_t mi = info->row_cut[i + 1] - info->row_cut[i];
    const int64_t k0 = info->k_cut[c0], kk = info->k_cut[c1] - k0;
    b200blas_step s = blank_step(B200BLAS_STEP_GEMM, d, SLOT_COMP);
    s.opa = opa; s.opb = opb;
    s.alpha[0] = alpha[0]; s.alpha[1] = alpha[1];
    if (uses_origin(d)) {
      s.x = a_origin(i, k0, kk); s.y = b_origin(j, s0, ns, k0, kk);
      s.z = region(-1, 2, info->row_cut[i], info->col_cut[j] + s0, mi, ns);
      s.beta[0] = first ? beta[0] : 1.0; s.beta[1] = first ? beta[1] : 0.0;
    } else {
      s.x = a_local(d, k0, kk); s.y = b_local(d, s0, ns, k0, kk);
      s.z = region(d, BUF_P, 0, s0, mi, ns);
      s.beta[0] = first ? 0.0 : 1.0; s.beta[1] = 0.0;
      // copies on one stream complete in order and chunks are issued in increasing order, so the delivery of the LAST
      // chunk of the range on each of the two inbound streams covers the whole range
      int wa = -1, wb = -1, wa2 = -1, wb2 = -1;
      for (int c = c0; c < c1; ++c) {
        const int sa = a_here[(size_t)d * nch + c], sb = b_here[(size_t)d * nch + c];
        if (sa >= 0) { if (sa < cap && steps[sa].stream == SLOT_FETCH) wa = sa; else wa2 = sa; }
        if (sb >= 0) { if (sb < cap && steps[sb].stream == SLOT_FETCH) wb = sb; else wb2 = sb; }
      }
      s.nwait = 4;
      s.wait[0] = std::max(wa, wb); s.wait[1] = std::max(wa2, wb2); s.wait[2] = -1; s.wait[3] = -1;
    }
    return B.add(s);
  };
  std::vector<int> cin_step(ndev, -1);
  for (int g = 0; g < ng; ++g) {
    const int c0 = info->group_first_chunk[g], c1 = info->group_first_chunk[g + 1];
    for (int c = c0; c < c1; ++c) emit_chunk(c);
    const bool last = g == ng - 1;
    if (last && host_origin && !beta_zero) {
      // the caller's C tiles travel last: they are only needed by the fold after the last product
      for (int d = 0; d < ndev; ++d) {
        if (tile_empty(d)) continue;
        const int i = d / pc, j = d % pc;
        const int64_t mi = info->row_cut[i + 1] - info->row_cut[i], nj = info->col_cut[j + 1] - info->col_cut[j];
        b200blas_step s = blank_step(B200BLAS_STEP_COPY, d, SLOT_FETCH);
        s.x = region(-1, 2, info->row_cut[i], info->col_cut[j], mi, nj);
        s.z = region(d, BUF_CIN, 0, 0, mi, nj);
        cin_step[d] = B.add(s);
      }
    }
    if (!last) {
      for (int d = 0; d < ndev; ++d) {
        if (tile_empty(d)) continue;
        const int j = d % pc;
        emit_gemm(d, 0, info->col_cut[j + 1] - info->col_cut[j], c0, c1, g == 0);
      }
      continue;
    }
    for (int sidx = 0; sidx < 3; ++sidx) {
      for (int d = 0; d < ndev; ++d) {
        if (tile_empty(d)) continue;
        const int i = d / pc, j = d % pc;
        const int64_t mi = info->row_cut[i + 1] - info->row_cut[i], nj = info->col_cut[j + 1] - info->col_cut[j];
        int64_t cuts[4];
        const int nsub = subslab_cuts(nj, cuts);
        if (sidx >= nsub) continue;
        const int64_t s0 = cuts[sidx], ns = cuts[sidx + 1] - s0;
        if (ns <= 0) continue;
        int prod = emit_gemm(d, s0, ns, c0, c1, g == 0);
        if (uses_origin(d)) continue;   // the root's tile was computed in place
        const b200blas_region csub = region(-1, 2, info->row_cut[i], info->col_cut[j] + s0, mi, ns);
        const b200blas_region psub = region(d, BUF_P, 0, s0, mi, ns);
        if (host_origin) {
          if (!beta_zero) {   // P := P + beta * C_in on the device, then one transfer back
            b200blas_step f = blank_step(B200BLAS_STEP_AXPBY, d, SLOT_COMP);
            f.x = region(d, BUF_CIN, 0, s0, mi, ns); f.z = psub;
            f.alpha[0] = beta[0]; f.alpha[1] = beta[1]; f.beta[0] = 1.0; f.beta[1] = 0.0;
            f.nwait = 1; f.wait[0] = cin_step[d];
            prod = B.add(f);
          }
          b200blas_step r = blank_step(B200BLAS_STEP_COPY, d, SLOT_RET);
          r.x = psub; r.z = csub;
          r.nwait = 1; r.wait[0] = prod;
          B.add(r);
        } else if (beta_zero) {   // straight into the root's C over NVLink
          b200blas_step r = blank_step(B200BLAS_STEP_COPY, d, SLOT_RET);
          r.x = psub; r.z = csub;
          r.nwait = 1; r.wait[0] = prod;
          B.add(r);
        } else {                  // into a staging tile on the root, folded there: C := beta * C + P
          b200blas_step r = blank_step(B200BLAS_STEP_COPY, d, SLOT_RET);
          r.x = psub; r.z = region(root, BUF_RECV0 + d, 0, s0, mi, ns);
          r.nwait = 1; r.wait[0] = prod;
          const int ret = B.add(r);
          b200blas_step f = blank_step(B200BLAS_STEP_AXPBY, root, SLOT_FOLD);
          f.x = r.z; f.z = csub;
          f.alpha[0] = 1.0; f.alpha[1] = 0.0; f.beta[0] = beta[0]; f.beta[1] = beta[1];
          f.nwait = 1; f.wait[0] = ret;
          B.add(f);
        }
      }
    }
  }
  info->nsteps = B.n;
  return B.n;
}

// ---- z := b * z + a * x on a rows x cols window (bandwidth-bound fold of beta*C) ----------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) axpby2d_kernel(T* __restrict__ z, int64_t ldz, const T* __restrict__ x, int64_t ldx,
                                                      int64_t rows, int64_t cols, T a, T b) {
  const int64_t total = rows * cols;
  for (int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * 256) {
    const int64_t i = idx % rows, j = idx / rows;
    z[i + j * ldz] = b * z[i + j * ldz] + a * x[i + j * ldx];
  }
}
template <typename R, typename T2>
__global__ void __launch_bounds__(256) axpby2d_cplx_kernel(T2* __restrict__ z, int64_t ldz, const T2* __restrict__ x, int64_t ldx,
                                                           int64_t rows, int64_t cols, R ar, R ai, R br, R bi) {
  const int64_t total = rows * cols;
  for (int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * 256) {
    const int64_t i = idx % rows, j = idx / rows;
    const T2 zv = z[i + j * ldz], xv = x[i + j * ldx];
    T2 out;
    out.x = br * zv.x - bi * zv.y + ar * xv.x - ai * xv.y;
    out.y = br * zv.y + bi * zv.x + ar * xv.y + ai * xv.x;
    z[i + j * ldz] = out;
  }
}

int launch_axpby(int type, void* z, int64_t ldz, const void* x, int64_t ldx, int64_t rows, int64_t cols, const double* a,
                 const double* b, cudaStream_t s) {
  if (rows <= 0 || cols <= 0) return 0;
  const int64_t total = rows * cols;
  const unsigned grid = (unsigned)std::min<int64_t>((total + 255) / 256, 148 * 16);
  switch (type) {
    case TY_S: axpby2d_kernel<float><<<grid, 256, 0, s>>>((float*)z, ldz, (const float*)x, ldx, rows, cols, (float)a[0], (float)b[0]); break;
    case TY_D: axpby2d_kernel<double><<<grid, 256, 0, s>>>((double*)z, ldz, (const double*)x, ldx, rows, cols, a[0], b[0]); break;
    case TY_C: axpby2d_cplx_kernel<float, float2><<<grid, 256, 0, s>>>((float2*)z, ldz, (const float2*)x, ldx, rows, cols, (float)a[0], (float)a[1], (float)b[0], (float)b[1]); break;
    default: axpby2d_cplx_kernel<double, double2><<<grid, 256, 0, s>>>((double2*)z, ldz, (const double2*)x, ldx, rows, cols, a[0], a[1], b[0], b[1]); break;
  }
  count_launch();
  return (int)cudaGetLastError();
}

// ---- executor ---------------------------------------------------------------------------------------------------------
struct DevCtx {
  int dev = -1;
  cudaStream_t st[NSLOT] = {};
  void* buf[BUF_RECV0 + MAXDEV] = {};
  size_t cap[BUF_RECV0 + MAXDEV] = {};
  std::vector<cudaEvent_t> ev;   // grow-only pool, handed out per call
  size_t ev_used = 0;
  cudaEvent_t join_ev[NSLOT] = {};
  PinnedRing ring_in, ring_out;
  bool ready = false;
  int init(int device) {
    if (ready && dev == device) return 0;
    dev = device;
    B200_CUDA_TRY(cudaSetDevice(dev));
    for (int i = 0; i < NSLOT; ++i) {
      // transfers outrank the product CTAs already queued on the device
      int lo = 0, hi = 0;
      cudaDeviceGetStreamPriorityRange(&lo, &hi);
      B200_CUDA_TRY(cudaStreamCreateWithPriority(&st[i], cudaStreamNonBlocking, i == SLOT_COMP ? lo : hi));
      B200_CUDA_TRY(cudaEventCreateWithFlags(&join_ev[i], cudaEventDisableTiming));
    }
    ready = true;
    return 0;
  }
  int reserve(int b, size_t bytes) {
    if (bytes <= cap[b]) return 0;
    B200_CUDA_TRY(cudaSetDevice(dev));
    if (buf[b]) { cudaFree(buf[b]); buf[b] = nullptr; cap[b] = 0; }
    B200_CUDA_TRY(cudaMalloc(&buf[b], bytes));
    cap[b] = bytes;
    return 0;
  }
  int event(cudaEvent_t* out) {
    if (ev_used == ev.size()) {
      cudaEvent_t e;
      B200_CUDA_TRY(cudaSetDevice(dev));
      B200_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      ev.push_back(e);
    }
    *out = ev[ev_used++];
    return 0;
  }
  void release() {
    if (dev < 0) return;
    cudaSetDevice(dev);
    for (auto& b : buf) { if (b) cudaFree(b); b = nullptr; }
    for (auto& c : cap) c = 0;
    for (auto& s : st) { if (s) cudaStreamDestroy(s); s = nullptr; }
    for (auto& e : join_ev) { if (e) cudaEventDestroy(e); e = nullptr; }
    for (auto e : ev) cudaEventDestroy(e);
    ev.clear(); ev_used = 0;
    ring_in.release(); ring_out.release();
    ready = false;
  }
};

struct Engine {
  std::mutex mu;                 // one multi-device product at a time per process
  DevCtx ctx[MAXDEV];
  cudaEvent_t start_ev = nullptr;
  int start_dev = -1;
  uint64_t peer_mask[64] = {};       // [CUDA device id] bit b: peer access to device b has been requested
  std::vector<b200blas_step> steps;
} g_engine;

int enable_peers(const int* devs, int n) {
  for (int a = 0; a < n; ++a)
    for (int b = 0; b < n; ++b) {
      if (devs[a] == devs[b] || devs[a] >= 64 || devs[b] >= 64) continue;
      uint64_t& mask = g_engine.peer_mask[devs[a]];
      if ((mask >> devs[b]) & 1ull) continue;
      int can = 0;
      cudaDeviceCanAccessPeer(&can, devs[a], devs[b]);
      if (can) {
        B200_CUDA_TRY(cudaSetDevice(devs[a]));
        const cudaError_t e = cudaDeviceEnablePeerAccess(devs[b], 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return (int)e;
        cudaGetLastError();
      }   // without peer access the copies still work (staged through the host by the driver), only slower
      mask |= 1ull << devs[b];
    }
  return 0;
}

struct Download {   // a finished C sub-slab waiting for its trip into pageable caller memory
  int dev_ix;
  cudaEvent_t ready;
  char* dst; size_t dpitch; const char* src; size_t spitch; size_t width; size_t ncols;
};

int execute(const GemmProblem& p, bool host_origin, cudaStream_t user_stream, uint64_t* h2d_bytes, uint64_t* d2h_bytes) {
  Engine& E = g_engine;
  std::lock_guard<std::mutex> lock(E.mu);
  const int ndev = devices_in_effect();
  int pr = 1, pc = 1;
  default_grid(ndev, &pr, &pc);
  int cur = 0;
  B200_CUDA_TRY(cudaGetDevice(&cur));
  // plan device 0 is the root: the device the operands live on (device-resident call) or the current device
  int devs[MAXDEV];
  {
    const int nvis = visible_devices();
    if (nvis < 1 || (nvis < ndev && !virtual_devices_ok())) return (int)cudaErrorInvalidDevice;
    int root = cur;
    if (!host_origin) {
      cudaPointerAttributes at;
      if (cudaPointerGetAttributes(&at, p.C) == cudaSuccess && at.type == cudaMemoryTypeDevice) root = at.device;
      cudaGetLastError();
    }
    for (int w = 0; w < ndev; ++w) devs[w] = (root + w) % nvis;   // distinct whenever nvis >= ndev
  }
  b200blas_plan_info info;
  E.steps.resize(4096);
  const int ns = build_plan(p.type, p.opa, p.opb, p.m, p.n, p.k, p.alpha, p.beta, ndev, pr, pc, host_origin, &info, E.steps.data(), (int)E.steps.size());
  if (ns < 0 || ns > (int)E.steps.size()) return (int)cudaErrorInvalidValue;
  const size_t es = (size_t)type_bytes(p.type);
  struct Restore { int dev; ~Restore() { cudaSetDevice(dev); } } restore{cur};
  B200_CUDA_TRY(enable_peers(devs, ndev));
  for (int d = 0; d < ndev; ++d) {
    DevCtx& c = E.ctx[d];
    if (c.ready && c.dev != devs[d]) c.release();
    B200_CUDA_TRY(c.init(devs[d]));
    c.ev_used = 0;
    for (int b = 0; b < 4; ++b) if (info.elems[d][b] > 0) B200_CUDA_TRY(c.reserve(b, (size_t)info.elems[d][b] * es));
  }
  if (!host_origin && !(p.beta[0] == 0.0 && p.beta[1] == 0.0))
    for (int d = 1; d < ndev; ++d) if (info.elems[d][BUF_P] > 0) B200_CUDA_TRY(E.ctx[0].reserve(BUF_RECV0 + d, (size_t)info.elems[d][BUF_P] * es));
  // everything waits for the work already queued on the caller's stream (device-resident call)
  if (!host_origin) {
    B200_CUDA_TRY(cudaSetDevice(devs[0]));
    if (!E.start_ev || E.start_dev != devs[0]) {
      if (E.start_ev) cudaEventDestroy(E.start_ev);
      B200_CUDA_TRY(cudaEventCreateWithFlags(&E.start_ev, cudaEventDisableTiming));
      E.start_dev = devs[0];
    }
    B200_CUDA_TRY(cudaEventRecord(E.start_ev, user_stream));
    for (int d = 0; d < ndev; ++d) {
      B200_CUDA_TRY(cudaSetDevice(devs[d]));
      for (int s = 0; s < NSLOT; ++s) B200_CUDA_TRY(cudaStreamWaitEvent(E.ctx[d].st[s], E.start_ev, 0));
    }
  }
  const void* origin[3] = {p.A, p.B, p.C};
  const int64_t origin_ld[3] = {p.lda, p.ldb, p.ldc};
  bool pageable[3] = {false, false, false};
  if (host_origin) for (int b = 0; b < 3; ++b) pageable[b] = is_pageable(origin[b]);
  auto resolve = [&](const b200blas_region& r, int64_t* ld) -> char* {
    if (r.loc < 0) { *ld = origin_ld[r.buf]; return (char*)origin[r.buf] + (size_t)(r.r0 + r.c0 * origin_ld[r.buf]) * es; }
    const int64_t l = r.buf >= BUF_RECV0 ? info.ld[r.buf - BUF_RECV0][BUF_P] : info.ld[r.loc][r.buf];
    *ld = l;
    return (char*)E.ctx[r.loc].buf[r.buf] + (size_t)(r.r0 + r.c0 * l) * es;
  };
  std::vector<cudaEvent_t> done(ns, nullptr);
  // pageable C: a downloader thread drains finished sub-slabs through the pinned rings while this thread keeps issuing
  std::deque<Download> dl_queue;
  std::mutex dl_mu;
  std::condition_variable dl_cv;
  bool dl_closed = false;
  int dl_err = 0;
  std::thread downloader;
  struct JoinGuard {
    std::thread& t; std::mutex& mu; std::condition_variable& cv; bool& closed;
    ~JoinGuard() { if (t.joinable()) { { std::lock_guard<std::mutex> l(mu); closed = true; } cv.notify_all(); t.join(); } }
  } join_guard{downloader, dl_mu, dl_cv, dl_closed};
  uint64_t h2d = 0, d2h = 0;
  int err = 0;
  for (int i = 0; i < ns && !err; ++i) {
    const b200blas_step& s = E.steps[i];
    DevCtx& c = E.ctx[s.dev];
    cudaStream_t stream = c.st[s.stream];
    if ((err = (int)cudaSetDevice(c.dev))) break;
    for (int w = 0; w < s.nwait && !err; ++w) err = (int)cudaStreamWaitEvent(stream, done[s.wait[w]], 0);
    if (err) break;
    if (s.kind == B200BLAS_STEP_COPY) {
      int64_t lds = 0, ldd = 0;
      const char* src = resolve(s.x, &lds);
      char* dst = resolve(s.z, &ldd);
      const size_t width = (size_t)s.x.rows * es, ncols = (size_t)s.x.cols;
      if (host_origin && s.x.loc < 0) {
        h2d += width * ncols;
        if (pageable[s.x.buf] && width * ncols >= ((size_t)1 << 20)) err = c.ring_in.h2d(dst, (size_t)ldd * es, src, (size_t)lds * es, width, ncols, stream);
        else err = (int)cudaMemcpy2DAsync(dst, (size_t)ldd * es, src, (size_t)lds * es, width, ncols, cudaMemcpyHostToDevice, stream);
      } else if (host_origin && s.z.loc < 0) {
        d2h += width * ncols;
        if (pageable[2] && width * ncols >= ((size_t)1 << 20)) {
          Download job;
          job.dev_ix = s.dev;
          if ((err = c.event(&job.ready))) break;
          if ((err = (int)cudaEventRecord(job.ready, stream))) break;
          job.dst = dst; job.dpitch = (size_t)ldd * es; job.src = src; job.spitch = (size_t)lds * es; job.width = width; job.ncols = ncols;
          if (!downloader.joinable()) {
            downloader = std::thread([&] {
              for (;;) {
                Download j;
                {
                  std::unique_lock<std::mutex> l(dl_mu);
                  dl_cv.wait(l, [&] { return !dl_queue.empty() || dl_closed; });
                  if (dl_queue.empty()) return;
                  j = dl_queue.front();
                  dl_queue.pop_front();
                }
                DevCtx& dc = E.ctx[j.dev_ix];
                int e = (int)cudaSetDevice(dc.dev);
                if (!e) e = (int)cudaEventSynchronize(j.ready);
                if (!e) e = dc.ring_out.d2h(j.dst, j.dpitch, j.src, j.spitch, j.width, j.ncols, dc.st[SLOT_RET]);
                if (e) { std::lock_guard<std::mutex> l(dl_mu); if (!dl_err) dl_err = e; }
              }
            });
          }
          { std::lock_guard<std::mutex> l(dl_mu); dl_queue.push_back(job); }
          dl_cv.notify_all();
        } else {
          err = (int)cudaMemcpy2DAsync(dst, (size_t)ldd * es, src, (size_t)lds * es, width, ncols, cudaMemcpyDeviceToHost, stream);
        }
      } else {
        err = (int)cudaMemcpy2DAsync(dst, (size_t)ldd * es, src, (size_t)lds * es, width, ncols, cudaMemcpyDefault, stream);
      }
    } else if (s.kind == B200BLAS_STEP_GEMM) {
      GemmProblem g;
      g.type = p.type; g.opa = s.opa; g.opb = s.opb;
      g.m = s.z.rows; g.n = s.z.cols; g.k = s.opa == OP_N ? s.x.cols : s.x.rows;
      g.alpha[0] = s.alpha[0]; g.alpha[1] = s.alpha[1]; g.beta[0] = s.beta[0]; g.beta[1] = s.beta[1];
      g.A = resolve(s.x, &g.lda); g.B = resolve(s.y, &g.ldb); g.C = resolve(s.z, &g.ldc);
      err = run_gemm_device(g, stream, B200BLAS_AUTO);
    } else {
      int64_t ldx = 0, ldz = 0;
      const char* x = resolve(s.x, &ldx);
      char* z = resolve(s.z, &ldz);
      err = launch_axpby(p.type, z, ldz, x, ldx, s.z.rows, s.z.cols, s.alpha, s.beta, stream);
    }
    if (!err && s.record) {
      err = c.event(&done[i]);
      if (!err) err = (int)cudaEventRecord(done[i], stream);
    }
  }
  if (downloader.joinable()) {
    { std::lock_guard<std::mutex> l(dl_mu); dl_closed = true; }
    dl_cv.notify_all();
    downloader.join();
    if (!err) err = dl_err;
  }
  if (err) {
    for (int d = 0; d < ndev; ++d) { cudaSetDevice(devs[d]); cudaDeviceSynchronize(); }
    return err;
  }
  if (host_origin) {
    for (int d = 0; d < ndev; ++d) {
      B200_CUDA_TRY(cudaSetDevice(devs[d]));
      for (int s = 0; s < NSLOT; ++s) B200_CUDA_TRY(cudaStreamSynchronize(E.ctx[d].st[s]));
    }
  } else {
    // the caller's stream continues after everything this call queued on any device (also protects the panel buffers
    // of the next call: it starts behind this join)
    for (int d = 0; d < ndev; ++d) {
      B200_CUDA_TRY(cudaSetDevice(devs[d]));
      for (int s = 0; s < NSLOT; ++s) B200_CUDA_TRY(cudaEventRecord(E.ctx[d].join_ev[s], E.ctx[d].st[s]));
    }
    B200_CUDA_TRY(cudaSetDevice(devs[0]));
    for (int d = 0; d < ndev; ++d)
      for (int s = 0; s < NSLOT; ++s) B200_CUDA_TRY(cudaStreamWaitEvent(user_stream, E.ctx[d].join_ev[s], 0));
  }
  if (h2d_bytes) *h2d_bytes = h2d;
  if (d2h_bytes) *d2h_bytes = d2h;
  return 0;
}

double min_flops() {   // read on every call (only reached when more than one device is enabled): tests lower it at run time
  const char* e = getenv("B200BLAS_MULTI_MIN_FLOPS");
  return e ? atof(e) : 4.6e11;
}

}  // namespace

bool multi_wanted(const GemmProblem& p) {
  if (devices_in_effect() <= 1) return false;
  if (p.uplo != UPLO_FULL || p.k <= 0 || p.m <= 0 || p.n <= 0) return false;
  const double flops = ((p.type == TY_C || p.type == TY_Z) ? 8.0 : 2.0) * (double)p.m * (double)p.n * (double)p.k;
  return flops >= min_flops();
}

int multi_gemm(const GemmProblem& p, bool host_origin, cudaStream_t stream, uint64_t* h2d_bytes, uint64_t* d2h_bytes) {
  note_variant("multi");   // overwritten by the kernels the devices run; kept if nothing launches
  return execute(p, host_origin, stream, h2d_bytes, d2h_bytes);
}

void multi_release() {
  std::lock_guard<std::mutex> lock(g_engine.mu);
  int cur = 0;
  cudaGetDevice(&cur);
  for (auto& c : g_engine.ctx) c.release();
  if (g_engine.start_ev) { cudaEventDestroy(g_engine.start_ev); g_engine.start_ev = nullptr; g_engine.start_dev = -1; }
  cudaSetDevice(cur);
  cudaGetLastError();
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200blas_set_devices(int n) {
  n = clamp_devices(n);
  g_ndev.store(n);
  return n;
}
int b200blas_get_devices(void) { return devices_in_effect(); }
int b200blas_set_grid(int pr, int pc) {
  if (pr == 0 && pc == 0) { g_grid_pr.store(0); g_grid_pc.store(0); return 0; }
  if (pr < 1 || pc < 1 || pr * pc > MAXDEV) return -1;
  g_grid_pr.store(pr); g_grid_pc.store(pc);
  return 0;
}
int b200blas_host_register(void* p, uint64_t bytes) { return (int)cudaHostRegister(p, (size_t)bytes, cudaHostRegisterPortable); }
int b200blas_host_unregister(void* p) { return (int)cudaHostUnregister(p); }

int b200blas_multi_plan(int type, char transa, char transb, int64_t m, int64_t n, int64_t k, const double* alpha2,
                        const double* beta2, int64_t lda, int64_t ldb, int64_t ldc, int ndev, int pr, int pc,
                        int host_origin, b200blas_plan_info* info, b200blas_step* steps, int cap) {
  (void)lda; (void)ldb; (void)ldc;   // the plan is in buffer coordinates; the leading dimensions only matter to the executor
  auto op = [](char x) { return (x == 'N' || x == 'n') ? OP_N : (x == 'T' || x == 't') ? OP_T : (x == 'C' || x == 'c') ? OP_C : -1; };
  const int opa = op(transa), opb = op(transb);
  if (type < 0 || type > 3 || opa < 0 || opb < 0 || !info || (cap > 0 && !steps)) return -1;
  if (pr == 0 && pc == 0) {
    pr = 1;
    for (int d = 1; d * d <= ndev; ++d) if (ndev % d == 0) pr = d;
    pc = ndev / pr;
  }
  std::vector<b200blas_step> tmp;
  if (cap <= 0) { tmp.resize(1); steps = tmp.data(); cap = 0; }
  return build_plan(type, opa, opb, m, n, k, alpha2, beta2, ndev, pr, pc, host_origin != 0, info, steps, cap);
}

}  // extern "C"
