#!/bin/bash
# usage: tools/gpu_multi.sh "8 4"   (world sizes to run on this box, in order)
mkdir -p gpurun_out
export TORCH_NCCL_SHOW_EAGER_INIT_P2P_SERIALIZATION_WARNING=0
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
port=29511
for N in $1; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port tools/dist_check.py > gpurun_out/dist_check_$N.log 2>&1
  port=$((port+1))
  grep "^dist" gpurun_out/dist_check_$N.log; tail -3 gpurun_out/dist_check_$N.log | grep -iE "error|assert" 
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_multi_$N.json 2> gpurun_out/bench_multi_$N.err
  port=$((port+1))
  tail -c 700 gpurun_out/bench_multi_$N.json; grep -iE "error|Traceback" gpurun_out/bench_multi_$N.err | head -5
done
