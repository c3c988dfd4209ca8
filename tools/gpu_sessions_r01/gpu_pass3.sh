#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu3.log 2>&1
tail -15 gpurun_out/pytest_gpu3.log
{
python -c "
import eigen_b200
print('peak tf32 tcgen05', eigen_b200.pipe_peak(3, 1000))
print('peak dmma', eigen_b200.pipe_peak(0, 500))
"
for cfg in A B C; do
  B200BLAS_DMMA_CFG=$cfg python tools/time_gemm.py d 8192 8192 8192 N N 5
  B200BLAS_DMMA_CFG=$cfg python tools/time_gemm.py d 16384 16384 256 N N 5
  B200BLAS_DMMA_CFG=$cfg python tools/time_gemm.py d 8192 8192 8192 T N 3
  B200BLAS_DMMA_CFG=$cfg python tools/time_gemm.py d 8192 8192 8192 N T 3
  B200BLAS_DMMA_CFG=$cfg python tools/time_gemm.py z 4096 4096 4096 N N 5
done
B200BLAS_DMMA_CFG=B python tools/time_gemm.py d 16384 16384 16384 N N 3
python tools/time_gemm.py s 8192 8192 8192 N N 5 tf32x3
python tools/time_gemm.py s 8192 8192 8192 T T 3 tf32x3
python tools/time_gemm.py s 8192 8192 8192 N N 3 simt
python tools/time_gemm.py s 4096 4096 4096 N N 5 tf32x3
} > gpurun_out/sweep3.log 2>&1
cat gpurun_out/sweep3.log
