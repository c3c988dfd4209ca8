#!/bin/bash
mkdir -p gpurun_out
{
for cfg in B D; do
  B200BLAS_DMMA_CFG=$cfg python tools/time_gemm.py d 8192 8192 8192 N N 5
  B200BLAS_DMMA_CFG=$cfg python tools/time_gemm.py d 16384 16384 256 N N 5
  B200BLAS_DMMA_CFG=$cfg python tools/time_gemm.py d 8192 8192 8192 T N 3
  B200BLAS_DMMA_CFG=$cfg python tools/time_gemm.py z 4096 4096 4096 N N 5
  B200BLAS_DMMA_CFG=$cfg python tools/time_gemm.py d 16384 16384 16384 N N 3
done
} > gpurun_out/sweep6.log 2>&1
cat gpurun_out/sweep6.log
B200BLAS_DMMA_CFG=D timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/pytest_gpu6_cfgD.log 2>&1
tail -3 gpurun_out/pytest_gpu6_cfgD.log
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu6.log 2>&1
tail -3 gpurun_out/pytest_gpu6.log
# DRAM traffic of the bench kernel at the bench size (one launch, --set full)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dmma_gemm_kernel -s 2 -c 1 -o gpurun_out/prof_r01_dmma_d_16384 \
  python tools/time_gemm.py d 16384 16384 16384 N N 1 > gpurun_out/ncu_full6.log 2>&1
ls -la gpurun_out/prof_r01_dmma_d_16384.ncu-rep
