#!/bin/bash
mkdir -p gpurun_out
{
B200BLAS_TF32_KCHUNK=100000 python tools/acc_sweep.py tf32x3
B200BLAS_TF32_KCHUNK=64 python tools/acc_sweep.py tf32x3
B200BLAS_TF32_KCHUNK=32 python tools/acc_sweep.py tf32x3
B200BLAS_TF32_KCHUNK=8 python tools/acc_sweep.py tf32x3
python tools/acc_sweep.py simt
for kc in 100000 64 32 16; do
  B200BLAS_TF32_KCHUNK=$kc python tools/time_gemm.py s 8192 8192 8192 N N 5 tf32x3
done
B200BLAS_TF32_KCHUNK=32 python tools/time_gemm.py c 4096 4096 4096 N N 5 tf32x3
python tools/time_gemm.py d 8192 8192 8192 N N 5
B200BLAS_DMMA_SYNC=mbar python tools/time_gemm.py d 8192 8192 8192 N N 5
python tools/time_gemm.py d 8192 8192 8192 T N 3
B200BLAS_DMMA_SYNC=mbar python tools/time_gemm.py d 8192 8192 8192 T N 3
python tools/time_gemm.py d 16384 16384 256 N N 5
B200BLAS_DMMA_SYNC=mbar python tools/time_gemm.py d 16384 16384 256 N N 5
python tools/time_gemm.py z 4096 4096 4096 N N 5
B200BLAS_DMMA_SYNC=mbar python tools/time_gemm.py z 4096 4096 4096 N N 5
python tools/time_gemm.py d 16384 16384 16384 N N 3
B200BLAS_DMMA_SYNC=mbar python tools/time_gemm.py d 16384 16384 16384 N N 3
python -c "
import eigen_b200
print('peak dmma', eigen_b200.pipe_peak(0, 1000))"
} > gpurun_out/sweep8.log 2>&1
cat gpurun_out/sweep8.log
B200BLAS_DMMA_SYNC=mbar timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/pytest_gpu8_mbar.log 2>&1
tail -3 gpurun_out/pytest_gpu8_mbar.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -m gpu > gpurun_out/pytest_gpu8.log 2>&1
tail -5 gpurun_out/pytest_gpu8.log
