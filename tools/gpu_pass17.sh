#!/bin/bash
mkdir -p gpurun_out
{
B200BLAS_DMMA_KSLICE=0 python tools/time_gemm.py d 8192 8192 8192 N N 5
python tools/time_gemm.py d 8192 8192 8192 T N 3
python tools/time_gemm.py d 16384 16384 256 N N 5
python tools/time_gemm.py z 4096 4096 4096 N N 5
python tools/time_gemm.py d 16384 16384 16384 N N 3
python tools/time_gemm.py d 2048 2048 2048 N N 8
} > gpurun_out/sweep17.log 2>&1
cat gpurun_out/sweep17.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -m gpu -k "d] or z] or c2 or c5 or lu_trailing or golden or committed" > gpurun_out/pytest_gpu17.log 2>&1
tail -3 gpurun_out/pytest_gpu17.log
