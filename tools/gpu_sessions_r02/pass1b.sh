#!/bin/bash
# round 2, GPU pass 1b (one B200): the driver bench line (kept), launch list of the same command, r2 drafts under their opt-ins
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tee gpurun_out/p1b_smi.txt
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/p1b_bench.json 2> gpurun_out/p1b_bench.err; echo "bench exit $?"; tail -3 gpurun_out/p1b_bench.err; cut -c1-1200 gpurun_out/p1b_bench.json
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/p1b_bench_ref.json 2> gpurun_out/p1b_bench_ref.err; echo "ref exit $?"; cut -c1-600 gpurun_out/p1b_bench_ref.json
echo "== drafts: correctness under the opt-in leaves"
for v in "B200BLAS_TRSM=inv" "B200BLAS_POTF2=cta" "B200BLAS_GETF2=cluster"; do
  env $v timeout 600 python -m pytest tests/test_gpu_level3.py tests/test_gpu_lapack.py tests/test_gpu_zz_golden_level3.py -x -q -m gpu > gpurun_out/p1b_draft_${v%%=*}.log 2>&1; echo "$v tests exit $?"; tail -6 gpurun_out/p1b_draft_${v%%=*}.log
done
echo "== drafts: timing"
for v in "X=0" "B200BLAS_TRSM=inv" "B200BLAS_POTF2=cta" "B200BLAS_GETF2=cluster" "B200BLAS_TRSM=inv B200BLAS_POTF2=cta B200BLAS_GETF2=cluster"; do
  for w in dtrsm8192 dpotrf8192 dgetrf8192 dpotrf16384 dgetrf16384; do
    env $v timeout 200 python bench.py --workload $w --steps 3 --warmup 3 2>/dev/null | tee -a gpurun_out/p1b_level3_lines.jsonl | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', d['metric'], round(d['value'],2), 'TF  ms', round(d['ms_per_step'],2), 'launches', d['roofline']['launches_per_step'], 'clk', d['clocks']['sm_mhz'])"
  done
done 2>&1 | tee gpurun_out/p1b_drafts_timing.txt
echo "== launch list of the bench command (shares only; never a bench value)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/p1b_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-configs > gpurun_out/p1b_ncu_bench.log 2>&1; echo "ncu exit $?"
