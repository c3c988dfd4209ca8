#!/bin/bash
# round-1 final validation on one B200: full GPU test suite, smoke, default bench line, level-3 / LAPACK measurement
mkdir -p gpurun_out
timeout 420 python -m pytest tests -x -q -m gpu > gpurun_out/final_tests.log 2>&1; echo "full gpu tests exit $?" | tee -a gpurun_out/final_tests.log
tail -4 gpurun_out/final_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke exit $?"; tail -8 gpurun_out/final_smoke.log
timeout 200 python bench.py --steps 5 --warmup 3 > gpurun_out/final_bench_dgemm16384.json 2> gpurun_out/final_bench.err; echo "bench exit $?"; cut -c1-600 gpurun_out/final_bench_dgemm16384.json
timeout 150 python tools/bench_level3.py --n 16384 --routines dsyrk,dtrsm,dpotrf,dgetrf > gpurun_out/final_level3_16384.jsonl 2> gpurun_out/final_level3_16384.err; echo "level3 16384 exit $?"
timeout 200 python tools/bench_level3.py --n 8192 > gpurun_out/final_level3_8192.jsonl 2> gpurun_out/final_level3_8192.err; echo "level3 8192 exit $?"
python - <<'PY'
import json
for f in ("gpurun_out/final_level3_16384.jsonl", "gpurun_out/final_level3_8192.jsonl"):
    try:
        for l in open(f):
            d = json.loads(l)
            print("%-28s value %7.2f  e2e %7.2f  frac %.3f  cpu %s  launches/step %d" % (d["metric"], d["value"], d["e2e"]["value"], d["roofline"]["frac"], ("%.4f" % d["cpu_baseline"]["value"]) if d["cpu_baseline"]["value"] else "-", d["roofline"]["launches_per_step"]))
    except Exception as e:
        print(f, e)
PY
