#!/bin/bash
# GPU pass 2: bench lines + ncu launch list + ncu full capture of the DMMA kernel
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_dgemm16384.json 2> gpurun_out/bench_dgemm16384.err
for w in dgemm_rankk zgemm4096 sgemm8192 cgemm4096; do
  timeout 600 python bench.py --steps 3 --warmup 3 --workload $w > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
# launch list (shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_dgemm8192.csv \
  python bench.py --steps 2 --warmup 3 --workload dgemm8192 > gpurun_out/ncu_launch.log 2>&1
# full capture of the top kernel
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dmma_gemm_kernel -s 3 -c 2 -o gpurun_out/prof_dmma_d \
  python bench.py --steps 2 --warmup 3 --workload dgemm8192 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/bench_*.json | cut -c1-1500
tail -5 gpurun_out/*.err
