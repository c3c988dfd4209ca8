#!/bin/bash
# first GPU pass: smoke, peaks, tests
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
lscpu | head -20 > gpurun_out/lscpu.txt; nproc >> gpurun_out/lscpu.txt
python -c "
import __graft_entry__ as g
g.smoke()
import eigen_b200
for p,n in ((0,'dmma'),(1,'dfma'),(2,'ffma')):
    print('peak', n, eigen_b200.pipe_peak(p, 1500))
" > gpurun_out/smoke.log 2>&1
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
tail -30 gpurun_out/smoke.log gpurun_out/pytest_gpu.log
