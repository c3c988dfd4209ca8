#!/bin/bash
# round 2, GPU pass 2 (TWO B200s): the multi-GPU partitioner behind dgemm_ on real peers
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv | tee gpurun_out/p2_smi.txt
nvidia-smi topo -m | head -8 | tee -a gpurun_out/p2_smi.txt
echo "== multi tests on real peers (plan devices beyond 2 share the two GPUs)"
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/p2_tests.log 2>&1; echo "tests exit $?"; tail -5 gpurun_out/p2_tests.log
echo "== bench N=2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/p2_bench_n2.json 2> gpurun_out/p2_bench_n2.err; echo "bench exit $?"; tail -5 gpurun_out/p2_bench_n2.err; cut -c1-3000 gpurun_out/p2_bench_n2.json
echo "== bench N=1 (same box, for the ratio)"
timeout 600 python bench.py --steps 10 --warmup 3 --no-configs > gpurun_out/p2_bench_n1.json 2> gpurun_out/p2_bench_n1.err; echo "bench exit $?"; cut -c1-400 gpurun_out/p2_bench_n1.json
echo "== bench_gemm -DHAVE_BLAS (the reference's own benchmark binary) with B200BLAS_NGPUS=2"
if [ -x oracle/_ref/bench_gemm_blas_d ]; then
  for n in 1 2; do B200BLAS_NGPUS=$n B200BLAS_LOG=1 timeout 300 oracle/_ref/bench_gemm_blas_d -s 8192 8192 8192 -t 3 2>&1 | tail -6; done | tee gpurun_out/p2_bench_gemm.txt
else ls oracle/_ref | head -30; fi
