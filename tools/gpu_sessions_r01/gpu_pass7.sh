#!/bin/bash
mkdir -p gpurun_out
{
python tools/time_gemm.py d 8192 8192 8192 N N 5
python tools/time_gemm.py d 8192 8192 8192 T N 3
python tools/time_gemm.py d 8192 8192 8192 N T 3
python tools/time_gemm.py d 8192 8192 8192 T T 3
python tools/time_gemm.py d 16384 16384 256 N N 5
python tools/time_gemm.py z 4096 4096 4096 N N 5
python tools/time_gemm.py z 4096 4096 4096 C T 3
python tools/time_gemm.py d 16384 16384 16384 N N 3
} > gpurun_out/sweep7.log 2>&1
cat gpurun_out/sweep7.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu7.log 2>&1
tail -5 gpurun_out/pytest_gpu7.log
# the reference's own benchmark binary with our library as its BLAS (pageable Eigen matrices)
(cd oracle/_ref && OMP_NUM_THREADS=$(nproc) timeout 600 ./bench_gemm_blas_d -s 4096 4096 4096 -t 3 > ../../gpurun_out/ref_bench_gemm_d_4096.txt 2>&1; OMP_NUM_THREADS=$(nproc) timeout 600 ./bench_gemm_blas_s -s 4096 4096 4096 -t 3 > ../../gpurun_out/ref_bench_gemm_s_4096.txt 2>&1)
cat gpurun_out/ref_bench_gemm_d_4096.txt gpurun_out/ref_bench_gemm_s_4096.txt
