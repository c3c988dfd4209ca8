#!/bin/bash
# usage: tools/gpu_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/dist_check_$N.log 2>&1
tail -12 gpurun_out/dist_check_$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_multi_$N.json 2> gpurun_out/bench_multi_$N.err
tail -c 1500 gpurun_out/bench_multi_$N.json; tail -5 gpurun_out/bench_multi_$N.err
