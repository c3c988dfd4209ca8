#!/bin/bash
mkdir -p gpurun_out
{
for kc in 4 8 16; do
  B200BLAS_TF32_KCHUNK=$kc python tools/time_gemm.py s 8192 8192 8192 N N 5 tf32x3
done
B200BLAS_TF32_KCHUNK=4 python tools/acc_sweep.py tf32x3
python tools/acc_sweep.py tf32x3
python tools/time_gemm.py s 16384 16384 16384 N N 3 tf32x3
python tools/time_gemm.py s 4096 4096 4096 T N 5 tf32x3
python tools/time_gemm.py s 2048 2048 2048 N N 8 tf32x3
python tools/time_gemm.py s 1024 1024 1024 N N 8 tf32x3
python tools/time_gemm.py s 512 512 512 N N 8 tf32x3
python tools/time_gemm.py s 512 512 512 N N 8 simt
python tools/time_gemm.py s 1024 1024 1024 N N 8 simt
python tools/time_gemm.py c 4096 4096 4096 N N 5 tf32x3
python tools/time_gemm.py c 2048 2048 2048 N N 5 tf32x3
python tools/time_gemm.py d 16384 16384 256 N N 5
python tools/time_gemm.py d 512 512 512 N N 8 simt
python tools/time_gemm.py d 512 512 512 N N 8 dmma
} > gpurun_out/sweep10.log 2>&1
cat gpurun_out/sweep10.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu10.log 2>&1
tail -5 gpurun_out/pytest_gpu10.log
