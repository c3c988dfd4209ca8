#!/bin/bash
# round 2, final GPU pass (one B200): what the driver will run -- full `pytest -m gpu`, smoke(), bench.py both arms -- plus the ncu evidence of
# the shipped dgemm launch (launch list of the bench command, --set full of one k-slice, DRAM bytes of one whole step)
mkdir -p gpurun_out
echo "== full GPU suite"; timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/p9_suite.log 2>&1; echo "suite exit $?"; tail -6 gpurun_out/p9_suite.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/p9_smoke.log 2>&1; echo "smoke exit $?"; tail -4 gpurun_out/p9_smoke.log
echo "== bench (ours)"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/p9_bench.json 2> gpurun_out/p9_bench.err; echo "bench exit $?"; tail -2 gpurun_out/p9_bench.err; cut -c1-700 gpurun_out/p9_bench.json
echo "== bench (reference arm)"; timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/p9_bench_ref.json 2>/dev/null; echo "ref exit $?"; cut -c1-300 gpurun_out/p9_bench_ref.json
echo "== potrf two-level blocking (opt-in): correctness + timing"
B200BLAS_POTRF_NB=256 timeout 600 python -m pytest tests/test_gpu_lapack.py -x -q -m gpu -k "potrf or failure or scale" > gpurun_out/p9_potrf_nb256.log 2>&1; echo "nb256 tests exit $?"; tail -3 gpurun_out/p9_potrf_nb256.log
for v in "X=0" "B200BLAS_POTRF_NB=256" "B200BLAS_POTRF_NB=512"; do
  for w in dpotrf8192 dpotrf16384 spotrf8192; do
    env $v timeout 200 python bench.py --workload $w --steps 3 --warmup 3 --no-configs 2>/dev/null | tee -a gpurun_out/p9_potrf_lines.jsonl | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', d['metric'], round(d['value'],2), 'TF  ms', round(d['ms_per_step'],2), 'launches', d['roofline']['launches_per_step'])"
  done
done 2>&1 | tee gpurun_out/p9_potrf_timing.txt
echo "== ncu: DRAM bytes + duration of the 8 k-slice launches of ONE timed step (warm-up 3 steps = 24 launches skipped)"
timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:dmma_gemm_kernel --launch-skip 24 -c 8 --csv --log-file gpurun_out/p9_dram_dgemm16384.csv python bench.py --steps 1 --warmup 3 --no-configs > /dev/null 2>&1; echo "ncu dram exit $?"
echo "== ncu --set full of one shipped k-slice launch"
timeout 400 ncu --set full --import-source on --clock-control none -k regex:dmma_gemm_kernel --launch-skip 26 -c 1 -o gpurun_out/p9_ncu_dmma_kslice python bench.py --steps 1 --warmup 3 --no-configs > /dev/null 2>&1; echo "ncu full exit $?"
echo "== ncu launch list of the bench command"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/p9_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-configs > /dev/null 2>&1; echo "ncu list exit $?"
