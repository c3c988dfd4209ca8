#!/bin/bash
mkdir -p gpurun_out
{
python tools/acc_sweep.py tf32x3
for kc in 8 16 32; do
  B200BLAS_TF32_KCHUNK=$kc python tools/time_gemm.py s 8192 8192 8192 N N 5 tf32x3
done
python tools/time_gemm.py s 2048 2048 2048 N N 8 tf32x3
python tools/time_gemm.py s 1024 1024 1024 N N 8 tf32x3
python tools/time_gemm.py s 16384 16384 16384 N N 3 tf32x3
python tools/time_gemm.py c 4096 4096 4096 N N 5 tf32x3
python tools/time_gemm.py d 8192 8192 8192 N N 5
python tools/time_gemm.py d 16384 16384 256 N N 5
python tools/time_gemm.py d 4096 4096 4096 N N 5
python tools/time_gemm.py d 2048 2048 2048 N N 8
python tools/time_gemm.py z 4096 4096 4096 N N 5
python tools/time_gemm.py d 16384 16384 16384 N N 3
} > gpurun_out/sweep9.log 2>&1
cat gpurun_out/sweep9.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu9.log 2>&1
tail -5 gpurun_out/pytest_gpu9.log
(cd oracle/_ref && export OMP_NUM_THREADS=$(nproc) && timeout 600 ./bench_gemm_blas_d -s 4096 4096 4096 -t 3 > ../../gpurun_out/ref_bench_gemm_d_4096.txt 2>&1; timeout 600 ./bench_gemm_blas_s -s 4096 4096 4096 -t 3 > ../../gpurun_out/ref_bench_gemm_s_4096.txt 2>&1; timeout 900 ./bench_gemm_blas_d -s 8192 8192 8192 -t 2 > ../../gpurun_out/ref_bench_gemm_d_8192.txt 2>&1)
grep -E "blas  real|eigen real|Warning|Matrix" gpurun_out/ref_bench_gemm_*.txt
