#!/bin/bash
# round 2, GPU pass 11 (one B200): the chained two-phase LU leaf (getf2_reg2_kernel) -- lapack tests + timing
mkdir -p gpurun_out
echo "== lapack tests + the reference's lu / cholesky tests"
timeout 900 python -m pytest tests/test_gpu_lapack.py tests/test_gpu_zz_golden_level3.py -x -q -m gpu > gpurun_out/p11_tests.log 2>&1; echo "tests exit $?"; tail -4 gpurun_out/p11_tests.log
timeout 600 python -m pytest tests/test_eigen_own_tests.py -x -q -m gpu -k "lu or cholesky" > gpurun_out/p11_eigen.log 2>&1; echo "eigen exit $?"; tail -3 gpurun_out/p11_eigen.log
echo "== timing"
for w in dpotrf8192 dgetrf8192 dpotrf16384 dgetrf16384 spotrf8192 sgetrf8192; do
  timeout 200 python bench.py --workload $w --steps 3 --warmup 3 --no-configs 2>/dev/null | tee -a gpurun_out/p11_level3_lines.jsonl | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['metric'], round(d['value'],2), 'TF  ms', round(d['ms_per_step'],2), 'launches', d['roofline']['launches_per_step'], 'clk', d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done 2>&1 | tee gpurun_out/p11_timing.txt
echo "== A/B: single leaves"
for w in dgetrf8192 dgetrf16384; do
  B200BLAS_GETF2_CHAIN=0 timeout 200 python bench.py --workload $w --steps 3 --warmup 3 --no-configs 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('CHAIN=0', d['metric'], round(d['value'],2), 'TF  ms', round(d['ms_per_step'],2), 'launches', d['roofline']['launches_per_step'])"
done 2>&1 | tee -a gpurun_out/p11_timing.txt
