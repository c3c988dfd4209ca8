"""The reference's OWN unit tests (test/product_{large,extra,small,trsolve,syrk}.cpp, cholesky.cpp, lu.cpp), compiled unmodified
with -DEIGEN_USE_BLAS by oracle/Makefile (target eigen_tests) and linked against libb200blas.so: every Eigen
`A*B` in them reaches our ?gemm_ through GeneralMatrixMatrix_BLAS.h:103, every triangular solve our ?trsm_
(TriangularSolverMatrix_BLAS.h:41-157) and every triangular product our ?trmm_ (TriangularMatrixMatrix_BLAS.h).
(test/product_symm.cpp does not compile in this snapshot with g++ 13 -- SelfAdjointView static assertion -- and
test/product_trmm.cpp binds no BLAS symbol, so neither is in the list.)  Needs a B200 (the binaries abort through
xerbla_ info=-1 otherwise: there is no CPU fallback)."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref")


def _run(name, variant, seed, ngpus=1):
    exe = os.path.join(BIN, "eigen_test_" + name)
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/eigen_test_%s not built (make -C oracle eigen_tests needs /root/reference)" % name)
    env = dict(os.environ)
    env.pop("B200BLAS_VARIANT", None)
    if variant != "auto":
        env["B200BLAS_VARIANT"] = variant
    if ngpus > 1:
        # the multi-GPU partitioner behind the same dgemm_ (include/b200blas.h section 3).  The products of these tests are
        # small, so the size threshold is lowered to route ALL of them through it; with fewer physical GPUs than plan
        # devices the plan devices share the GPUs (B200BLAS_MULTI_VIRTUAL), which runs the same executor.
        env.update(B200BLAS_NGPUS=str(ngpus), B200BLAS_MULTI_MIN_FLOPS="0", B200BLAS_MULTI_VIRTUAL="1")
    # r<repeat> s<seed>: test/main.h:136,766-780
    p = subprocess.run([exe, "r3", "s%d" % seed], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=900, text=True)
    assert p.returncode == 0, p.stdout[-3000:]
    assert "ERROR #-1" not in p.stdout, "a CUDA failure was reported through xerbla_"


@pytest.mark.parametrize("variant", ["auto", "dmma", "tf32x3"])
@pytest.mark.parametrize("name", ["product_large", "product_extra", "product_small"])
def test_eigen_product_tests_pass_on_the_gpu_library(name, variant):
    _run(name, variant, 12345)


@pytest.mark.parametrize("name", ["product_trsolve", "cholesky", "lu", "product_syrk"])
def test_eigen_solver_tests_pass_on_the_gpu_library(name):
    """Binaries import ?trsm_/?trmm_/?gemm_ from libb200blas.so (nm -D); ?gemv_/?trmv_ come from the reference blas."""
    _run(name, "auto", 4242)


@pytest.mark.parametrize("name", ["product_large", "product_extra", "lu"])
def test_eigen_tests_pass_with_the_multi_gpu_partitioner(name):
    """the reference's own tests, unmodified, with B200BLAS_NGPUS=8 -- every product is cut into the
    2x4 grid inside dgemm_ (Parallelizer.h:85-157 is what the reference does at this point)."""
    _run(name, "auto", 777, ngpus=8)


def test_eigen_tensor_contraction_test_runs_on_the_gpu_library():
    """the reference's own unsupported/test/cxx11_tensor_contract_cuda.cu (GpuDevice contractions, ColMajor and
    RowMajor, the m / k / n size sweeps and the scalar case), compiled unmodified against the Tensor headers with the
    five-line binding of INTEGRATION.md section 4 (oracle/patch_tensor_contraction.py).  The test compares every result with
    Eigen's CPU contraction itself; B200BLAS_LOG shows that the products really ran on our kernels."""
    exe = os.path.join(BIN, "eigen_test_tensor_contract_cuda")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/eigen_test_tensor_contract_cuda not built (make -C oracle tensor_test needs /root/reference)")
    env = dict(os.environ, B200BLAS_LOG="1")
    p = subprocess.run([exe, "r1", "s99"], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=900, text=True)
    assert p.returncode == 0, p.stdout[-3000:]
    assert "mismatch detected" not in p.stdout
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("[b200blas] contract")]
    assert len(lines) >= 50, "the contractions did not go through b200blas_contract_dev:\n" + p.stdout[-2000:]
    assert not any("FAILED" in ln for ln in lines)
    assert any("tf32x3" in ln for ln in lines) and any("simt" in ln for ln in lines), "both the tensor and the SIMT variant should appear"
