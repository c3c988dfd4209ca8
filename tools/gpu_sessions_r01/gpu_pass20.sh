#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lapack.py tests/test_gpu_level3.py -x -q -k "potrf or getrf or failure or scale or tri_" > gpurun_out/p20_tests.log 2>&1; echo "tests exit $?" | tee -a gpurun_out/p20_tests.log
tail -6 gpurun_out/p20_tests.log
timeout 300 python tools/time_lapack.py 8192 16384 > gpurun_out/p20_time_lapack.log 2>&1; tail -12 gpurun_out/p20_time_lapack.log
timeout 300 python tools/time_level3.py 8192 2>&1 | grep -E "trsm|trmm" > gpurun_out/p20_time_level3.log; cat gpurun_out/p20_time_level3.log
