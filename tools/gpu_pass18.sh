#!/bin/bash
mkdir -p gpurun_out
{
for d in 0 1 2; do
B200BLAS_DMMA_KSLICE=0 B200BLAS_DMMA_DIAG=$d python tools/time_gemm.py d 8192 8192 8192 N N 4
done
B200BLAS_DMMA_KSLICE=0 B200BLAS_DMMA_DIAG=2 python tools/time_gemm.py d 8192 8192 8192 T N 4
B200BLAS_DMMA_KSLICE=0 B200BLAS_DMMA_DIAG=0 python tools/time_gemm.py d 8192 8192 8192 T N 4
} > gpurun_out/sweep18.log 2>&1
cat gpurun_out/sweep18.log
