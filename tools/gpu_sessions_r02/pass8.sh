#!/bin/bash
# round 2, GPU pass 8 (one B200): leaf kernels after the ncu-driven fixes (REDUX arg-max, fast reciprocal, roots off the loop, batched loads),
# the Tensor contraction seam, timing, launch lists
mkdir -p gpurun_out
echo "== small cases"; timeout 600 python tools/sanitize_small.py > gpurun_out/p8_small.log 2>&1; echo "small exit $?"; tail -2 gpurun_out/p8_small.log
echo "== tensor contraction test (reference's own, patched header copy)"
B200BLAS_LOG=1 timeout 600 oracle/_ref/eigen_test_tensor_contract_cuda r1 s99 > gpurun_out/p8_tensor.log 2>&1; echo "tensor exit $?"; grep -c "b200blas\] contract" gpurun_out/p8_tensor.log; grep "contract" gpurun_out/p8_tensor.log | awk '{print $NF}' | sort | uniq -c | head; tail -3 gpurun_out/p8_tensor.log
echo "== lapack / level3 / eigen tests"
timeout 1500 python -m pytest tests/test_gpu_lapack.py tests/test_gpu_level3.py tests/test_gpu_zz_golden_level3.py tests/test_eigen_own_tests.py -x -q -m gpu > gpurun_out/p8_tests.log 2>&1; echo "tests exit $?"; tail -5 gpurun_out/p8_tests.log
echo "== timing"
for v in "X=0" "B200BLAS_LOOKAHEAD=0"; do
  for w in dtrsm8192 dpotrf8192 dgetrf8192 dpotrf16384 dgetrf16384 spotrf8192 sgetrf8192; do
    env $v timeout 200 python bench.py --workload $w --steps 3 --warmup 3 --no-configs 2>/dev/null | tee -a gpurun_out/p8_level3_lines.jsonl | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', d['metric'], round(d['value'],2), 'TF  ms', round(d['ms_per_step'],2), 'launches', d['roofline']['launches_per_step'], 'clk', d['clocks']['sm_mhz'], d['clocks']['reasons'])"
  done
done 2>&1 | tee gpurun_out/p8_timing.txt
echo "== ncu --set full"
timeout 300 ncu --set full --import-source on --clock-control none -k regex:getf2_reg_kernel --launch-skip 40 -c 1 -o gpurun_out/p8_ncu_getf2_reg python bench.py --workload dgetrf8192 --steps 1 --warmup 3 > /dev/null 2>&1; echo "ncu getf2 exit $?"
timeout 300 ncu --set full --import-source on --clock-control none -k regex:potf2_block_kernel --launch-skip 40 -c 1 -o gpurun_out/p8_ncu_potf2_block python bench.py --workload dpotrf8192 --steps 1 --warmup 3 > /dev/null 2>&1; echo "ncu potf2 exit $?"
timeout 300 ncu --set full --import-source on --clock-control none -k regex:tri_block_solve_kernel --launch-skip 40 -c 1 -o gpurun_out/p8_ncu_block_solve python bench.py --workload dpotrf8192 --steps 1 --warmup 3 > /dev/null 2>&1; echo "ncu block_solve exit $?"
echo "== launch lists (first 1500 launches)"
for w in dpotrf8192 dgetrf8192; do
  timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/p8_launches_$w.csv python bench.py --workload $w --steps 1 --warmup 1 > /dev/null 2>&1; echo "ncu $w exit $?"
done
