#!/bin/bash
mkdir -p gpurun_out
{
python -c "
import eigen_b200
for p,n in ((0,'dmma 32 warps/SM'),(10,'dmma 16 warps/SM'),(11,'dmma 8 warps/SM'),(12,'dmma 4 warps/SM')):
    print('peak', n, eigen_b200.pipe_peak(p, 500))"
B200BLAS_DMMA_KSLICE=0 python tools/time_gemm.py d 8192 8192 8192 N N 5
B200BLAS_DMMA_KSLICE=0 B200BLAS_DMMA_DIAG=1 python tools/time_gemm.py d 8192 8192 8192 N N 5
B200BLAS_DMMA_KSLICE=0 B200BLAS_DMMA_SYNC=bar python tools/time_gemm.py d 8192 8192 8192 N N 5
} > gpurun_out/sweep16.log 2>&1
cat gpurun_out/sweep16.log
