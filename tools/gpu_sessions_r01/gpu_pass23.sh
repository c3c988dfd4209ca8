#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/p23_tests.log 2>&1; echo "full gpu tests exit $?" | tee -a gpurun_out/p23_tests.log
tail -8 gpurun_out/p23_tests.log
timeout 300 python tools/time_lapack.py 8192 16384 > gpurun_out/p23_time_lapack.log 2>&1; tail -12 gpurun_out/p23_time_lapack.log
timeout 300 python tools/time_level3.py 8192 2>&1 | grep -E "^d|^z" > gpurun_out/p23_time_level3.log; cat gpurun_out/p23_time_level3.log
for sz in 256 512 768 1024 1536; do
  for tile in small big; do echo -n "dgemm $sz^3 tile=$tile: "; B200BLAS_DMMA_TILE=$tile timeout 100 python tools/time_gemm.py d $sz $sz $sz N N 20 2>&1 | tail -1; done
done
