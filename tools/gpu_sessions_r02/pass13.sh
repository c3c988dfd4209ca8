#!/bin/bash
# round 2, GPU pass 13 (one B200): final defaults (chained LU leaf for float only) -- lapack tests, smoke, timing lines that go into DESIGN
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_lapack.py tests/test_gpu_zz_golden_level3.py -x -q -m gpu > gpurun_out/p13_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/p13_tests.log
timeout 300 python -m pytest tests/test_eigen_own_tests.py -x -q -m gpu -k "lu or cholesky or trsolve" > gpurun_out/p13_eigen.log 2>&1; echo "eigen exit $?"; tail -2 gpurun_out/p13_eigen.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/p13_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/p13_smoke.log
for w in dtrsm8192 dpotrf8192 dgetrf8192 dpotrf16384 dgetrf16384 spotrf8192 sgetrf8192 spotrf16384 sgetrf16384; do
  timeout 200 python bench.py --workload $w --steps 3 --warmup 3 --no-configs 2>/dev/null | tee -a gpurun_out/p13_level3_lines.jsonl | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['metric'], round(d['value'],2), 'TF  ms', round(d['ms_per_step'],2), 'launches', d['roofline']['launches_per_step'], 'clk', d['clocks']['sm_mhz'], d['clocks']['reasons'], 'e2e', round(d['e2e']['value'],2))"
done 2>&1 | tee gpurun_out/p13_timing.txt
