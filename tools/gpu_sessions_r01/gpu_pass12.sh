#!/bin/bash
mkdir -p gpurun_out
{
timeout 120 python tools/time_gemm.py c 4096 4096 4096 N N 5 tf32x3
B200BLAS_TF32_PAIR=0 timeout 120 python tools/time_gemm.py c 4096 4096 4096 N N 5 tf32x3
timeout 120 python tools/time_gemm.py c 4096 4096 4096 C T 3 tf32x3
timeout 120 python tools/time_gemm.py c 8192 8192 8192 N N 3 tf32x3
timeout 120 python tools/time_gemm.py c 2048 2048 2048 N N 5 tf32x3
timeout 120 python tools/time_gemm.py s 8192 8192 8192 N N 5
timeout 120 python tools/time_gemm.py s 1024 1024 1024 N N 8
B200BLAS_TF32_PAIR=1 timeout 120 python tools/time_gemm.py s 1024 1024 1024 N N 8 tf32x3
} > gpurun_out/sweep12.log 2>&1
cat gpurun_out/sweep12.log
B200BLAS_TF32_PAIR=1 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "s] or c] or device_pointer" > gpurun_out/pytest_gpu12_pair.log 2>&1
tail -4 gpurun_out/pytest_gpu12_pair.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu12.log 2>&1
tail -4 gpurun_out/pytest_gpu12.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench12_dgemm16384.json 2> gpurun_out/bench12.err
python bench.py --steps 5 --warmup 3 --workload sgemm8192 > gpurun_out/bench12_sgemm8192.json 2>> gpurun_out/bench12.err
tail -c 1200 gpurun_out/bench12_dgemm16384.json; tail -3 gpurun_out/bench12.err
