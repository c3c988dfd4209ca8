#!/bin/bash
# round 2, GPU pass 4 (one B200): the new blocked / look-ahead factorizations -- correctness first, then timing against the forced old variants
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== small cases (also the sanitizer script, plain)"; timeout 600 python tools/sanitize_small.py > gpurun_out/p4_small.log 2>&1; echo "small exit $?"; tail -12 gpurun_out/p4_small.log
echo "== lapack / level3 / golden / eigen's own tests"
timeout 1200 python -m pytest tests/test_gpu_lapack.py tests/test_gpu_level3.py tests/test_gpu_zz_golden_level3.py tests/test_eigen_own_tests.py -x -q -m gpu > gpurun_out/p4_tests.log 2>&1; echo "tests exit $?"; tail -8 gpurun_out/p4_tests.log
echo "== same with look-ahead off"
B200BLAS_LOOKAHEAD=0 timeout 900 python -m pytest tests/test_gpu_lapack.py -x -q -m gpu > gpurun_out/p4_tests_nolook.log 2>&1; echo "tests exit $?"; tail -4 gpurun_out/p4_tests_nolook.log
echo "== timing"
for v in "X=0" "B200BLAS_LOOKAHEAD=0" "B200BLAS_POTRF=rec B200BLAS_GETRF=rec B200BLAS_GETF2=slab B200BLAS_TRSM=subst"; do
  for w in dtrsm8192 dpotrf8192 dgetrf8192 dpotrf16384 dgetrf16384 spotrf8192 sgetrf8192; do
    env $v timeout 200 python bench.py --workload $w --steps 3 --warmup 3 2>/dev/null | tee -a gpurun_out/p4_level3_lines.jsonl | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', d['metric'], round(d['value'],2), 'TF  ms', round(d['ms_per_step'],2), 'launches', d['roofline']['launches_per_step'], 'clk', d['clocks']['sm_mhz'], d['clocks']['reasons'])"
  done
done 2>&1 | tee gpurun_out/p4_timing.txt
echo "== launch lists"
for w in dpotrf8192 dgetrf8192; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/p4_launches_$w.csv python bench.py --workload $w --steps 1 --warmup 3 > /dev/null 2>&1; echo "ncu $w exit $?"
done
