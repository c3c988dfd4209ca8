#!/bin/bash
# potrf / getrf: parity tests + device-resident timing; then the reference's own solver tests through the library
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lapack.py -x -q > gpurun_out/p18_lapack_tests.log 2>&1; echo "lapack tests exit $?" | tee -a gpurun_out/p18_lapack_tests.log
tail -25 gpurun_out/p18_lapack_tests.log
timeout 300 python tools/time_lapack.py 8192 16384 > gpurun_out/p18_time_lapack.log 2>&1; tail -12 gpurun_out/p18_time_lapack.log
timeout 900 python -m pytest tests/test_eigen_own_tests.py -x -q -k solver > gpurun_out/p18_eigen_solver_tests.log 2>&1; echo "eigen solver tests exit $?"; tail -8 gpurun_out/p18_eigen_solver_tests.log
