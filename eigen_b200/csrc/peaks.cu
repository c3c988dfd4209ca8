// eigen_b200/csrc/peaks.cu -- pipe-peak micro-benchmarks (roofline denominators measured on the box).
//
// MEASURED_PEAKS.json carries HBM and bf16 numbers only; the dgemm / sgemm kernels are bound by the FP64 DMMA pipe
// and the TF32 tensor pipe, so those peaks are measured here with register-resident loops (no memory traffic):
//   pipe 0  mma.sync.aligned.m8n8k4.f64 (SASS DMMA.8x8x4), 16 independent accumulator tiles per warp
//   pipe 1  DFMA, 16 independent chains per thread
//   pipe 2  FFMA, 16 independent chains per thread
//   pipe 3  tcgen05.mma kind::tf32 M=128 N=256 K=8 from shared-memory descriptors (see gemm_tf32x3.cu)
#include "common.cuh"

namespace b200 {
double tf32_pipe_peak(int millis);  // gemm_tf32x3.cu

namespace {

__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters) {
  double acc[16][2];
#pragma unroll
  for (int i = 0; i < 16; ++i) { acc[i][0] = 0.0; acc[i][1] = 0.0; }
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(acc[i][0]), "+d"(acc[i][1]) : "d"(a), "d"(b));
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i][0] + acc[i][1];
  if (s == 123.456) out[0] = s;
}

__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters) {
  double acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-3 + i;
  const double a = 1.0000001, b = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  if (s == 123.456) out[0] = s;
}

__global__ void __launch_bounds__(256) ffma_peak_kernel(float* out, int iters) {
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-3f + i;
  const float a = 1.0000001f, b = 1e-9f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], a, b);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  if (s == 123.456f) out[0] = s;
}

template <typename F>
double time_loop(F launch, double flops_per_launch, int millis) {
  cudaEvent_t e0, e1;
  if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) return -1.0;
  launch();  // warm-up
  if (cudaDeviceSynchronize() != cudaSuccess) return -1.0;
  double best = 0.0, total_ms = 0.0;
  while (total_ms < millis) {
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) return -1.0;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    total_ms += ms;
    const double tf = flops_per_launch / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return best;
}

}  // namespace

double pipe_peak(int pipe, int millis) {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1.0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  void* out = nullptr;
  if (cudaMalloc(&out, 64) != cudaSuccess) return -1.0;
  const int blocks = sms * 4, threads = 256;  // 32 warps / SM
  double r = -1.0;
  if (pipe == 0) {
    const int iters = 4096;
    const double flops = (double)blocks * (threads / 32) * iters * 16.0 * 512.0;  // 8x8x4 MAC = 512 flop
    r = time_loop([&] { dmma_peak_kernel<<<blocks, threads>>>((double*)out, iters); count_launch(); }, flops, millis);
  } else if (pipe == 1) {
    const int iters = 4096;
    const double flops = (double)blocks * threads * iters * 16.0 * 2.0;
    r = time_loop([&] { dfma_peak_kernel<<<blocks, threads>>>((double*)out, iters); count_launch(); }, flops, millis);
  } else if (pipe == 2) {
    const int iters = 16384;
    const double flops = (double)blocks * threads * iters * 16.0 * 2.0;
    r = time_loop([&] { ffma_peak_kernel<<<blocks, threads>>>((float*)out, iters); count_launch(); }, flops, millis);
  } else if (pipe == 3) {
    r = tf32_pipe_peak(millis);
  } else if (pipe == 10 || pipe == 11 || pipe == 12) {
    // occupancy sensitivity of the DMMA pipe: 16 / 8 / 4 warps per SM (4 / 2 / 1 per sub-partition), 16 independent tiles each
    const int iters = 4096;
    const int nb = pipe == 10 ? sms * 2 : pipe == 11 ? sms : sms;
    const int nt = pipe == 12 ? 128 : 256;
    const double flops = (double)nb * (nt / 32) * iters * 16.0 * 512.0;
    r = time_loop([&] { dmma_peak_kernel<<<nb, nt>>>((double*)out, iters); count_launch(); }, flops, millis);
  }
  cudaFree(out);
  return r;
}

}  // namespace b200
