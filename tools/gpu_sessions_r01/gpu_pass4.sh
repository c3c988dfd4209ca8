#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu4.log 2>&1
tail -15 gpurun_out/pytest_gpu4.log
{
python tools/time_gemm.py s 8192 8192 8192 N N 8 tf32x3
python tools/time_gemm.py c 4096 4096 4096 N N 8 tf32x3
python tools/time_gemm.py c 4096 4096 4096 C T 4 tf32x3
python tools/time_gemm.py c 4096 4096 4096 N N 3 simt
python tools/time_gemm.py s 16384 16384 16384 N N 3 tf32x3
python tools/time_gemm.py s 2048 2048 2048 N N 8 tf32x3
python tools/time_gemm.py s 2048 2048 2048 N N 8 simt
python tools/time_gemm.py s 1024 1024 1024 N N 8 tf32x3
python tools/time_gemm.py s 1024 1024 1024 N N 8 simt
python tools/time_gemm.py s 512 512 512 N N 8 tf32x3
python tools/time_gemm.py s 512 512 512 N N 8 simt
python tools/time_gemm.py d 512 512 512 N N 8 dmma
python tools/time_gemm.py d 512 512 512 N N 8 simt
python tools/time_gemm.py d 256 256 256 N N 8 dmma
python tools/time_gemm.py d 256 256 256 N N 8 simt
python tools/time_gemm.py d 128 128 128 N N 8 dmma
python tools/time_gemm.py d 128 128 128 N N 8 simt
python tools/time_gemm.py d 2400 24 24 N N 8 dmma
python tools/time_gemm.py d 2400 24 24 N N 8 simt
python tools/time_gemm.py d 24 2400 2400 N N 8 dmma
python tools/time_gemm.py d 24 2400 2400 N N 8 simt
} > gpurun_out/sweep4.log 2>&1
cat gpurun_out/sweep4.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench4_dgemm16384.json 2> gpurun_out/bench4_dgemm16384.err
tail -c 3000 gpurun_out/bench4_dgemm16384.json; tail -3 gpurun_out/bench4_dgemm16384.err
# ncu: launch list for the bench command and full captures of the two tensor kernels
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r01_bench_dgemm8192.csv \
  python bench.py --steps 2 --warmup 3 --workload dgemm8192 > gpurun_out/ncu_launch4.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dmma_gemm_kernel -s 2 -c 1 -o gpurun_out/prof_r01_dmma_d_cfgB \
  python tools/time_gemm.py d 8192 8192 8192 N N 2 > gpurun_out/ncu_full4a.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tf32x3_gemm_kernel -s 2 -c 1 -o gpurun_out/prof_r01_tf32x3_s \
  python tools/time_gemm.py s 8192 8192 8192 N N 2 tf32x3 > gpurun_out/ncu_full4b.log 2>&1
ls -la gpurun_out/*.ncu-rep
