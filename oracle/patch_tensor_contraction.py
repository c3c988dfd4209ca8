#!/usr/bin/env python
"""Test infrastructure: builds a PATCHED COPY of the reference's Tensor module header tree (never the reference itself, never
committed) in which the GpuDevice contraction evaluator offers its product to libb200blas first.

    python oracle/patch_tensor_contraction.py /root/reference oracle/_ref/eigen_patched

The patch is what INTEGRATION.md section 4 asks a maintainer to add to
unsupported/Eigen/CXX11/src/Tensor/TensorContractionCuda.h (evalTyped, :1350-1390): one #include and one guarded call.
Everything else in the copy is the unmodified reference (unsupported/ is copied, Eigen/ is a symlink)."""
import os
import shutil
import sys

ref, out = sys.argv[1], sys.argv[2]
if os.path.isdir(out):
    shutil.rmtree(out)
os.makedirs(out)
os.symlink(os.path.join(ref, "Eigen"), os.path.join(out, "Eigen"))
shutil.copytree(os.path.join(ref, "unsupported"), os.path.join(out, "unsupported"), ignore=shutil.ignore_patterns("test", "doc", "bench"))
path = os.path.join(out, "unsupported/Eigen/CXX11/src/Tensor/TensorContractionCuda.h")
src = open(path).read()

INCLUDE_ANCHOR = "namespace Eigen {"
INCLUDE = "#ifdef EIGEN_USE_B200BLAS\n#include <b200blas_eigen_tensor.h>\n#endif\n\n"
CALL_ANCHOR = "    // zero out the result buffer (which must be of size at least m * n * sizeof(Scalar)\n"
CALL = """#ifdef EIGEN_USE_B200BLAS
    // sm_100a GEMM kernels for every contraction that is a matrix product in memory; otherwise the kernels below
    if (b200blas_eigen::try_gemm<Scalar, LhsScalar, RhsScalar>(this->m_leftImpl.data(), this->m_rightImpl.data(), buffer, m, n, k,
            this->m_left_nocontract_strides, this->m_i_strides, this->m_left_contracting_strides, this->m_right_contracting_strides,
            this->m_k_strides, this->m_right_nocontract_strides, this->m_j_strides, (void*)this->m_device.stream()))
      return;
#endif
"""
assert src.count(INCLUDE_ANCHOR) >= 1 and src.count(CALL_ANCHOR) == 1, "the reference header changed: update the anchors"
src = src.replace(INCLUDE_ANCHOR, INCLUDE + INCLUDE_ANCHOR, 1)
src = src.replace(CALL_ANCHOR, CALL + CALL_ANCHOR, 1)
open(path, "w").write(src)
print("patched", path)

# Build fix, unrelated to the path: test/main.h:72 poisons isnan() to catch unprotected calls, and
# SpecialFunctionsImpl.h:538,731 of this reference revision has two of them (numext::isnan(x)), so the unmodified test does not
# compile with any compiler as soon as the Tensor module is included.  The copy protects the two calls the way the rest of Eigen does.
sf = os.path.join(out, "unsupported/Eigen/src/SpecialFunctions/SpecialFunctionsImpl.h")
txt = open(sf).read()
txt = txt.replace("numext::isnan(", "(numext::isnan)(")
open(sf, "w").write(txt)
print("protected isnan calls in", sf)
