#!/bin/bash
mkdir -p gpurun_out
export B200BLAS_TF32_PAIR=1
{
timeout 120 python tools/acc_sweep.py tf32x3
timeout 120 python tools/time_gemm.py s 8192 8192 8192 N N 5 tf32x3
timeout 120 python tools/time_gemm.py s 4096 4096 4096 T N 5 tf32x3
timeout 120 python tools/time_gemm.py s 2048 2048 2048 N N 8 tf32x3
timeout 120 python tools/time_gemm.py s 16384 16384 16384 N N 3 tf32x3
B200BLAS_TF32_PAIR=0 timeout 120 python tools/time_gemm.py s 8192 8192 8192 N N 5 tf32x3
} > gpurun_out/sweep11.log 2>&1
cat gpurun_out/sweep11.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "s-" > gpurun_out/pytest_gpu11_pair.log 2>&1
tail -5 gpurun_out/pytest_gpu11_pair.log
timeout 300 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k "c3" > gpurun_out/pytest_gpu11_pair_full.log 2>&1
tail -3 gpurun_out/pytest_gpu11_pair_full.log
