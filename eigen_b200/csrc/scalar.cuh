// eigen_b200/csrc/scalar.cuh -- scalar traits of the four BLAS types for the SIMT kernels (real and interleaved complex).
// conj() is the device-side conj_if / conj_helper of the reference (Eigen/src/Core/util/BlasUtil.h:43-124).
#pragma once
#include "common.cuh"

namespace b200 {

template <typename T> struct Sc;  // scalar traits
template <> struct Sc<float> {
  using real = float;
  static __device__ __forceinline__ float zero() { return 0.f; }
  static __device__ __forceinline__ float make(double re, double) { return (float)re; }
  static __device__ __forceinline__ float conj(float a) { return a; }
  static __device__ __forceinline__ void fma(float& c, float a, float b) { c = fmaf(a, b, c); }
  static __device__ __forceinline__ float mul(float a, float b) { return a * b; }
  static __device__ __forceinline__ bool is_zero(float a) { return a == 0.f; }
};
template <> struct Sc<double> {
  using real = double;
  static __device__ __forceinline__ double zero() { return 0.0; }
  static __device__ __forceinline__ double make(double re, double) { return re; }
  static __device__ __forceinline__ double conj(double a) { return a; }
  static __device__ __forceinline__ void fma(double& c, double a, double b) { c = ::fma(a, b, c); }
  static __device__ __forceinline__ double mul(double a, double b) { return a * b; }
  static __device__ __forceinline__ bool is_zero(double a) { return a == 0.0; }
};
template <> struct Sc<float2> {
  using real = float;
  static __device__ __forceinline__ float2 zero() { return make_float2(0.f, 0.f); }
  static __device__ __forceinline__ float2 make(double re, double im) { return make_float2((float)re, (float)im); }
  static __device__ __forceinline__ float2 conj(float2 a) { return make_float2(a.x, -a.y); }
  static __device__ __forceinline__ void fma(float2& c, float2 a, float2 b) {
    c.x = fmaf(a.x, b.x, c.x); c.x = fmaf(-a.y, b.y, c.x);
    c.y = fmaf(a.x, b.y, c.y); c.y = fmaf(a.y, b.x, c.y);
  }
  static __device__ __forceinline__ float2 mul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
  }
  static __device__ __forceinline__ bool is_zero(float2 a) { return a.x == 0.f && a.y == 0.f; }
};
template <> struct Sc<double2> {
  using real = double;
  static __device__ __forceinline__ double2 zero() { return make_double2(0.0, 0.0); }
  static __device__ __forceinline__ double2 make(double re, double im) { return make_double2(re, im); }
  static __device__ __forceinline__ double2 conj(double2 a) { return make_double2(a.x, -a.y); }
  static __device__ __forceinline__ void fma(double2& c, double2 a, double2 b) {
    c.x = ::fma(a.x, b.x, c.x); c.x = ::fma(-a.y, b.y, c.x);
    c.y = ::fma(a.x, b.y, c.y); c.y = ::fma(a.y, b.x, c.y);
  }
  static __device__ __forceinline__ double2 mul(double2 a, double2 b) {
    return make_double2(::fma(a.x, b.x, -a.y * b.y), ::fma(a.x, b.y, a.y * b.x));
  }
  static __device__ __forceinline__ bool is_zero(double2 a) { return a.x == 0.0 && a.y == 0.0; }
};

// additional helpers used by the triangular kernels
template <typename T> __device__ __forceinline__ T sc_recip(T a);
template <> __device__ __forceinline__ float sc_recip<float>(float a) { return 1.f / a; }
template <> __device__ __forceinline__ double sc_recip<double>(double a) { return 1.0 / a; }
template <> __device__ __forceinline__ float2 sc_recip<float2>(float2 a) {   // Smith's algorithm: no spurious overflow
  if (fabsf(a.x) >= fabsf(a.y)) { const float r = a.y / a.x, d = a.x + a.y * r; return make_float2(1.f / d, -r / d); }
  const float r = a.x / a.y, d = a.x * r + a.y; return make_float2(r / d, -1.f / d);
}
template <> __device__ __forceinline__ double2 sc_recip<double2>(double2 a) {
  if (fabs(a.x) >= fabs(a.y)) { const double r = a.y / a.x, d = a.x + a.y * r; return make_double2(1.0 / d, -r / d); }
  const double r = a.x / a.y, d = a.x * r + a.y; return make_double2(r / d, -1.0 / d);
}
// Reciprocal for latency-critical pivots (the leaf kernels of ?potrf_ / ?getrf_ wait on it once per column): the IEEE double
// division is a ~30-instruction dependent sequence (350+ cycles of "wait" stalls in ncu); a float MUFU.RCP seed refined by two
// Newton steps in double is ~8 dependent instructions and accurate to about one ulp.  Arguments outside the float range
// (and 0, Inf, NaN) take the IEEE path.
__device__ __forceinline__ double fast_rcp(double x) {
  const float xf = (float)x;
  if (!(fabsf(xf) > 1e-30f && fabsf(xf) < 1e30f)) return 1.0 / x;
  double r = (double)__frcp_rn(xf);
  double e = ::fma(-x, r, 1.0);
  r = ::fma(r, e, r);
  e = ::fma(-x, r, 1.0);
  return ::fma(r, e, r);
}
__device__ __forceinline__ float fast_rcp(float x) { return __frcp_rn(x); }
template <typename T> __device__ __forceinline__ T sc_fast_recip(T a);
template <> __device__ __forceinline__ float sc_fast_recip<float>(float a) { return fast_rcp(a); }
template <> __device__ __forceinline__ double sc_fast_recip<double>(double a) { return fast_rcp(a); }
template <> __device__ __forceinline__ float2 sc_fast_recip<float2>(float2 a) {   // Smith's algorithm: no spurious overflow
  if (fabsf(a.x) >= fabsf(a.y)) { const float r = a.y * fast_rcp(a.x), d = fast_rcp(a.x + a.y * r); return make_float2(d, -r * d); }
  const float r = a.x * fast_rcp(a.y), d = fast_rcp(a.x * r + a.y); return make_float2(r * d, -d);
}
template <> __device__ __forceinline__ double2 sc_fast_recip<double2>(double2 a) {
  if (fabs(a.x) >= fabs(a.y)) { const double r = a.y * fast_rcp(a.x), d = fast_rcp(a.x + a.y * r); return make_double2(d, -r * d); }
  const double r = a.x * fast_rcp(a.y), d = fast_rcp(a.x * r + a.y); return make_double2(r * d, -d);
}
template <typename T> __device__ __forceinline__ T sc_one() { return Sc<T>::make(1.0, 0.0); }
// c -= a * b
template <typename T> __device__ __forceinline__ void sc_fnma(T& c, T a, T b);
template <> __device__ __forceinline__ void sc_fnma<float>(float& c, float a, float b) { c = fmaf(-a, b, c); }
template <> __device__ __forceinline__ void sc_fnma<double>(double& c, double a, double b) { c = ::fma(-a, b, c); }
template <> __device__ __forceinline__ void sc_fnma<float2>(float2& c, float2 a, float2 b) {
  c.x = fmaf(-a.x, b.x, c.x); c.x = fmaf(a.y, b.y, c.x);
  c.y = fmaf(-a.x, b.y, c.y); c.y = fmaf(-a.y, b.x, c.y);
}
template <> __device__ __forceinline__ void sc_fnma<double2>(double2& c, double2 a, double2 b) {
  c.x = ::fma(-a.x, b.x, c.x); c.x = ::fma(a.y, b.y, c.x);
  c.y = ::fma(-a.x, b.y, c.y); c.y = ::fma(-a.y, b.x, c.y);
}

}  // namespace b200
