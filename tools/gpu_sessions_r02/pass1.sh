#!/bin/bash
# round 2, GPU pass 1 (one B200): new parity tests + multi-GPU executor on virtual devices, full suite, bench line, r2 drafts
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tee gpurun_out/p1_smi.txt
echo "== new tests first"; timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/p1_newtests.log 2>&1; echo "new tests exit $?"; tail -25 gpurun_out/p1_newtests.log
echo "== rest of the suite"; timeout 900 python -m pytest tests -q -m gpu --deselect tests/test_gpu_multi.py --deselect tests/test_gpu_parity.py > gpurun_out/p1_suite.log 2>&1; echo "suite exit $?"; tail -15 gpurun_out/p1_suite.log
echo "== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/p1_smoke.log 2>&1; echo "smoke exit $?"; tail -12 gpurun_out/p1_smoke.log
echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/p1_bench.json 2> gpurun_out/p1_bench.err; echo "bench exit $?"; tail -3 gpurun_out/p1_bench.err; cut -c1-1500 gpurun_out/p1_bench.json
echo "== drafts: correctness under the opt-in leaves"
for v in "B200BLAS_TRSM=inv" "B200BLAS_POTF2=cta" "B200BLAS_GETF2=cluster"; do
  env $v timeout 600 python -m pytest tests/test_gpu_level3.py tests/test_gpu_lapack.py tests/test_gpu_zz_golden_level3.py -x -q -m gpu > gpurun_out/p1_draft_${v%%=*}.log 2>&1; echo "$v tests exit $?"; tail -6 gpurun_out/p1_draft_${v%%=*}.log
done
echo "== drafts: timing"
for v in "X=0" "B200BLAS_TRSM=inv" "B200BLAS_POTF2=cta" "B200BLAS_GETF2=cluster"; do
  for w in dtrsm8192 dpotrf8192 dgetrf8192; do
    env $v timeout 200 python bench.py --workload $w --steps 3 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', d['metric'], round(d['value'],2), 'TF  ms', round(d['ms_per_step'],2), 'launches', d['roofline']['launches_per_step'])"
  done
done 2>&1 | tee gpurun_out/p1_drafts_timing.txt
