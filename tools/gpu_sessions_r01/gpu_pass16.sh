#!/bin/bash
# rank-k updates: parity tests + device-resident timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rankk.py -x -q > gpurun_out/p16_rankk_tests.log 2>&1; echo "rankk tests exit $?" | tee -a gpurun_out/p16_rankk_tests.log
tail -15 gpurun_out/p16_rankk_tests.log
timeout 300 python tools/time_rankk.py 16384 16384 > gpurun_out/p16_time_rankk.log 2>&1; cat gpurun_out/p16_time_rankk.log
