#!/bin/bash
mkdir -p gpurun_out
export TORCH_NCCL_SHOW_EAGER_INIT_P2P_SERIALIZATION_WARNING=0
{
B200BLAS_TF32_TILE=128 python tools/time_gemm.py s 8192 8192 8192 N N 8 tf32x3
B200BLAS_TF32_TILE=256 python tools/time_gemm.py s 8192 8192 8192 N N 8 tf32x3
B200BLAS_TF32_TILE=256 python tools/time_gemm.py s 16384 16384 16384 N N 4 tf32x3
B200BLAS_TF32_TILE=256 python tools/time_gemm.py s 4096 4096 4096 N T 8 tf32x3
B200BLAS_TF32_TILE=128 python tools/time_gemm.py s 4096 4096 4096 N T 8 tf32x3
B200BLAS_TF32_TILE=256 python tools/time_gemm.py s 2048 2048 2048 N N 8 tf32x3
} > gpurun_out/sweep5.log 2>&1
cat gpurun_out/sweep5.log
B200BLAS_TF32_TILE=256 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "random_shapes or device_pointer or xblat3" > gpurun_out/pytest_gpu5.log 2>&1
tail -5 gpurun_out/pytest_gpu5.log
./tools/gpu_multi.sh "2"
