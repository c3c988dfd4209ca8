#!/bin/bash
# round 2, GPU pass 10 (one B200): interchanges without a permutation pass (perm_apply_direct_kernel) and the new potrf default -- lapack tests + timing
mkdir -p gpurun_out
echo "== lapack tests + the reference's lu / cholesky tests"
timeout 900 python -m pytest tests/test_gpu_lapack.py tests/test_gpu_zz_golden_level3.py -x -q -m gpu > gpurun_out/p10_tests.log 2>&1; echo "tests exit $?"; tail -4 gpurun_out/p10_tests.log
timeout 600 python -m pytest tests/test_eigen_own_tests.py -x -q -m gpu -k "lu or cholesky" > gpurun_out/p10_eigen.log 2>&1; echo "eigen exit $?"; tail -3 gpurun_out/p10_eigen.log
echo "== timing"
for w in dpotrf8192 dgetrf8192 dpotrf16384 dgetrf16384 spotrf8192 sgetrf8192; do
  timeout 200 python bench.py --workload $w --steps 3 --warmup 3 --no-configs 2>/dev/null | tee -a gpurun_out/p10_level3_lines.jsonl | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['metric'], round(d['value'],2), 'TF  ms', round(d['ms_per_step'],2), 'launches', d['roofline']['launches_per_step'], 'clk', d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done 2>&1 | tee gpurun_out/p10_timing.txt
