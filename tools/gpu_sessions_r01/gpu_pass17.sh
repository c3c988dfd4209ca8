#!/bin/bash
# remaining level-3 routines: parity tests + device-resident timing
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_level3.py -x -q > gpurun_out/p17_level3_tests.log 2>&1; echo "level3 tests exit $?" | tee -a gpurun_out/p17_level3_tests.log
tail -25 gpurun_out/p17_level3_tests.log
timeout 300 python tools/time_level3.py 8192 > gpurun_out/p17_time_level3.log 2>&1; cat gpurun_out/p17_time_level3.log | tail -30
