#!/bin/bash
mkdir -p gpurun_out
{
for ks in 0 4096 2048; do
  B200BLAS_DMMA_KSLICE=$ks python tools/time_gemm.py d 16384 16384 16384 N N 3
done
for ks in 0 2048; do
  B200BLAS_DMMA_KSLICE=$ks python tools/time_gemm.py z 8192 8192 8192 N N 3
done
python tools/time_gemm.py d 8192 8192 8192 N N 3
} > gpurun_out/sweep14.log 2>&1
cat gpurun_out/sweep14.log
for ks in 4096 2048; do
B200BLAS_DMMA_KSLICE=$ks timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none -k regex:dmma_gemm_kernel -s 8 -c 8 --csv --log-file gpurun_out/ncu14_kslice$ks.csv \
  python tools/time_gemm.py d 16384 16384 16384 N N 1 > gpurun_out/ncu14.log 2>&1
done
grep -h "dmma" gpurun_out/ncu14_kslice*.csv | cut -d, -f5,13-15 | cut -c1-200 | head -70
timeout 600 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_parity.py -x -q -m gpu -k "d] or z] or c2 or c5 or lu_trailing" > gpurun_out/pytest_gpu14.log 2>&1
tail -3 gpurun_out/pytest_gpu14.log
