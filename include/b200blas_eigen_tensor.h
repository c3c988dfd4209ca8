/* include/b200blas_eigen_tensor.h -- C++ glue between Eigen's Tensor GpuDevice contraction evaluator and libb200blas.so.
 *
 * Reference seam: TensorEvaluator<const TensorContractionOp<...>, GpuDevice>::evalTyped
 * (unsupported/Eigen/CXX11/src/Tensor/TensorContractionCuda.h:1320-1390) builds LhsMapper / RhsMapper over the operands'
 * strides and launches EigenContractionKernel / EigenFloatContractionKernel.  With -DEIGEN_USE_B200BLAS the evaluator first
 * offers the product to this header (the five-line binding is shown in INTEGRATION.md section 5); contractions whose three index groups (left free, contracted, right free) are each contiguous-mergeable --
 * every matrix product and every contraction of leading / trailing index groups -- run on the sm_100a GEMM kernels, anything
 * else falls through to the reference's own kernels unchanged.  Header only; needs nothing but b200blas.h. */
#ifndef B200BLAS_EIGEN_TENSOR_H
#define B200BLAS_EIGEN_TENSOR_H

#include <complex>
#include <cstddef>
#include <cstdint>

#include "b200blas.h"

namespace b200blas_eigen {

template <typename Scalar> struct type_code { static const int value = -1; };
template <> struct type_code<float> { static const int value = B200BLAS_S; };
template <> struct type_code<double> { static const int value = B200BLAS_D; };
template <> struct type_code<std::complex<float> > { static const int value = B200BLAS_C; };
template <> struct type_code<std::complex<double> > { static const int value = B200BLAS_Z; };

/* A group of tensor indices with memory strides `mem` and logical (cumulative-product) strides `logical` addresses memory like
 * ONE index of stride mem[0] iff mem[d] == mem[0] * logical[d] for every d. */
template <typename Strides>
inline bool mergeable(const Strides& mem, const Strides& logical) {
  for (std::size_t d = 1; d < mem.size(); ++d)
    if ((int64_t)mem[d] != (int64_t)mem[0] * (int64_t)logical[d]) return false;
  return true;
}
template <typename Strides>
inline int64_t base_stride(const Strides& mem) { return mem.size() > 0 ? (int64_t)mem[0] : 1; }

/* true: the product has been queued on `stream` and `out` (m x n, column-major) will hold lhs * rhs. */
template <typename Scalar, typename LhsScalar, typename RhsScalar, typename Index, typename LeftFree, typename RightFree, typename Contract>
inline bool try_gemm(const LhsScalar* lhs, const RhsScalar* rhs, Scalar* out, Index m, Index n, Index k,
                     const LeftFree& left_free_mem, const LeftFree& left_free_logical, const Contract& left_contract_mem,
                     const Contract& right_contract_mem, const Contract& contract_logical, const RightFree& right_free_mem,
                     const RightFree& right_free_logical, void* stream) {
  const int type = type_code<Scalar>::value;
  if (type < 0 || type_code<LhsScalar>::value != type || type_code<RhsScalar>::value != type || !lhs || !rhs) return false;
  if (!mergeable(left_free_mem, left_free_logical) || !mergeable(right_free_mem, right_free_logical) ||
      !mergeable(left_contract_mem, contract_logical) || !mergeable(right_contract_mem, contract_logical))
    return false;
  return b200blas_contract_dev(type, (int64_t)m, (int64_t)n, (int64_t)k, lhs, base_stride(left_free_mem), base_stride(left_contract_mem),
                               rhs, base_stride(right_contract_mem), base_stride(right_free_mem), out, (int64_t)m, stream) == 0;
}

}  /* namespace b200blas_eigen */
#endif
