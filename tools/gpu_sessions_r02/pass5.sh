#!/bin/bash
# round 2, GPU pass 5 (one B200): ncu --set full of the two new leaf kernels, the DMMA tail split, the e2e head pipeline, sanitizer
mkdir -p gpurun_out
echo "== dmma tail split: parity + C1"
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "dmma or random_shapes or huge" > gpurun_out/p5_tests.log 2>&1; echo "tests exit $?"; tail -4 gpurun_out/p5_tests.log
for v in "X=0" "B200BLAS_DMMA_TAIL=0"; do
  for w in dgemm2048 dgemm8192; do
    env $v timeout 200 python bench.py --workload $w --steps 20 --warmup 5 --no-configs 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', d['metric'], round(d['value'],2), 'TF frac', round(d['roofline']['frac'],3), d['kernel'], 'e2e', round(d['e2e']['value'],2), round(d['e2e_pinned']['value'],2))"
  done
done 2>&1 | tee gpurun_out/p5_tail.txt
echo "== e2e probe (head pipeline), copy-thread sweep"
for ct in default 12 16; do
  if [ $ct = default ]; then timeout 300 python tools/e2e_probe.py --steps 2; else B200BLAS_COPY_THREADS=$ct timeout 300 python tools/e2e_probe.py --steps 2; fi
done 2>/dev/null | tee gpurun_out/p5_e2e.jsonl
echo "== ncu --set full: LU register panel and Cholesky block"
timeout 400 ncu --set full --import-source on --clock-control none -k regex:getf2_reg_kernel --launch-skip 40 -c 1 -o gpurun_out/p5_ncu_getf2_reg python bench.py --workload dgetrf8192 --steps 1 --warmup 3 > /dev/null 2>&1; echo "ncu getf2 exit $?"
timeout 400 ncu --set full --import-source on --clock-control none -k regex:potf2_block_kernel --launch-skip 40 -c 1 -o gpurun_out/p5_ncu_potf2_block python bench.py --workload dpotrf8192 --steps 1 --warmup 3 > /dev/null 2>&1; echo "ncu potf2 exit $?"
ls -la gpurun_out/*.ncu-rep
echo "== sanitizer (bounded)"
timeout 420 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_small.py > gpurun_out/p5_memcheck.log 2>&1; echo "memcheck exit $?"; tail -4 gpurun_out/p5_memcheck.log
timeout 420 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_small.py > gpurun_out/p5_racecheck.log 2>&1; echo "racecheck exit $?"; tail -4 gpurun_out/p5_racecheck.log
