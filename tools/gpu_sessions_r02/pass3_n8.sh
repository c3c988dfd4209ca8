#!/bin/bash
# round 2, GPU pass 3 (EIGHT B200s): grid A/B table through the C ABI, then the driver's bench line at N=8
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv | tee gpurun_out/p3_smi.txt
echo "== grid A/B (one process, all devices)"
timeout 300 python tools/scale_grid.py --ndev 8 --grids 2x4,1x8,4x2 --steps 5 2> gpurun_out/p3_grid.err | tee gpurun_out/p3_grid_n8.jsonl
timeout 200 python tools/scale_grid.py --ndev 4 --grids 2x2,1x4 --steps 5 2>> gpurun_out/p3_grid.err | tee gpurun_out/p3_grid_n4.jsonl
tail -3 gpurun_out/p3_grid.err
echo "== multi tests on 8 real peers"
timeout 300 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/p3_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/p3_tests.log
echo "== bench N=8"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/p3_bench_n8.json 2> gpurun_out/p3_bench_n8.err; echo "bench exit $?"; tail -5 gpurun_out/p3_bench_n8.err; cut -c1-3500 gpurun_out/p3_bench_n8.json
