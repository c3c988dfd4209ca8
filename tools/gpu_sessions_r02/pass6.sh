#!/bin/bash
# round 2, GPU pass 6 (one B200): compact (rotated-row) leaf kernels -- correctness, timing, ncu; racecheck of the DMMA loaders under both sync modes
mkdir -p gpurun_out
echo "== small cases"; timeout 600 python tools/sanitize_small.py > gpurun_out/p6_small.log 2>&1; echo "small exit $?"; tail -3 gpurun_out/p6_small.log
echo "== lapack tests"
timeout 1200 python -m pytest tests/test_gpu_lapack.py tests/test_gpu_zz_golden_level3.py tests/test_eigen_own_tests.py -x -q -m gpu > gpurun_out/p6_tests.log 2>&1; echo "tests exit $?"; tail -5 gpurun_out/p6_tests.log
echo "== timing"
for v in "X=0" "B200BLAS_LOOKAHEAD=0"; do
  for w in dpotrf8192 dgetrf8192 dpotrf16384 dgetrf16384 spotrf8192 sgetrf8192; do
    env $v timeout 200 python bench.py --workload $w --steps 3 --warmup 3 2>/dev/null | tee -a gpurun_out/p6_level3_lines.jsonl | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', d['metric'], round(d['value'],2), 'TF  ms', round(d['ms_per_step'],2), 'launches', d['roofline']['launches_per_step'], 'clk', d['clocks']['sm_mhz'], d['clocks']['reasons'])"
  done
done 2>&1 | tee gpurun_out/p6_timing.txt
echo "== ncu --set full"
timeout 300 ncu --set full --import-source on --clock-control none -k regex:getf2_reg_kernel --launch-skip 40 -c 1 -o gpurun_out/p6_ncu_getf2_reg python bench.py --workload dgetrf8192 --steps 1 --warmup 3 > /dev/null 2>&1; echo "ncu getf2 exit $?"
timeout 300 ncu --set full --import-source on --clock-control none -k regex:potf2_block_kernel --launch-skip 40 -c 1 -o gpurun_out/p6_ncu_potf2_block python bench.py --workload dpotrf8192 --steps 1 --warmup 3 > /dev/null 2>&1; echo "ncu potf2 exit $?"
echo "== launch lists"
for w in dpotrf8192 dgetrf8192; do
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/p6_launches_$w.csv python bench.py --workload $w --steps 1 --warmup 3 > /dev/null 2>&1; echo "ncu $w exit $?"
done
echo "== racecheck of the DMMA loaders: __syncthreads ring vs mbarrier ring"
cat > /tmp/rc.py <<'PY'
import sys; sys.path.insert(0, "tools"); import runpy
PY
B200BLAS_DMMA_SYNC=bar timeout 300 compute-sanitizer --tool racecheck --print-limit 5 python -c "
import sys, os
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, eigen_b200, oracle_api as oa
rng = np.random.default_rng(1)
for (t, ta, tb, m, n, k) in (('d','N','N',300,200,200), ('d','T','C',1400,1300,100), ('z','C','N',150,130,90)):
    A = oa.rand_matrix(rng, t, *((m, k) if ta == 'N' else (k, m))); B = oa.rand_matrix(rng, t, *((k, n) if tb == 'N' else (n, k)))
    c = oa.rand_matrix(rng, t, m, n)
    assert eigen_b200.gemm_host(t, ta, tb, m, n, k, 0.7, A, A.shape[0], B, B.shape[0], 1.3, c, m) == 0
    print('ok', t, m, n, k, eigen_b200.last_variant())
" > gpurun_out/p6_racecheck_dmma_bar.log 2>&1; echo "racecheck(bar) exit $?"; tail -4 gpurun_out/p6_racecheck_dmma_bar.log
timeout 300 compute-sanitizer --tool synccheck --print-limit 10 python tools/sanitize_small.py > gpurun_out/p6_synccheck.log 2>&1; echo "synccheck exit $?"; tail -3 gpurun_out/p6_synccheck.log
