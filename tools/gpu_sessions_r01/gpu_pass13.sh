#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "concurrent or pageable" > gpurun_out/pytest_gpu13.log 2>&1
tail -5 gpurun_out/pytest_gpu13.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dmma_gemm_kernel -s 2 -c 1 -o gpurun_out/prof_r01_dmma_d_final \
  python tools/time_gemm.py d 8192 8192 8192 N N 1 > gpurun_out/ncu13a.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tf32x3_gemm_pair_kernel -s 0 -c 1 -o gpurun_out/prof_r01_tf32x3_pair \
  python tools/time_gemm.py s 8192 8192 8192 N N 1 > gpurun_out/ncu13b.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dmma_gemm_kernel -s 2 -c 1 -o gpurun_out/prof_r01_dmma_d_16384_final \
  python tools/time_gemm.py d 16384 16384 16384 N N 1 > gpurun_out/ncu13c.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r01_bench_dgemm16384.csv \
  python bench.py --steps 2 --warmup 3 > gpurun_out/ncu13_launch.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_r01_bench_sgemm8192.csv \
  python bench.py --steps 2 --warmup 3 --workload sgemm8192 > gpurun_out/ncu13_launch_s.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_r01_bench_*
