#!/bin/bash
mkdir -p gpurun_out
timeout 500 python tools/bench_level3.py --n 8192 > gpurun_out/p21_level3_8192.jsonl 2> gpurun_out/p21_level3_8192.err; echo "exit $?"; tail -3 gpurun_out/p21_level3_8192.err
timeout 500 python tools/bench_level3.py --n 16384 --routines dsyrk,dtrsm,dpotrf,dgetrf > gpurun_out/p21_level3_16384.jsonl 2> gpurun_out/p21_level3_16384.err; echo "exit $?"; tail -3 gpurun_out/p21_level3_16384.err
python - <<'PY'
import json
for f in ("gpurun_out/p21_level3_8192.jsonl", "gpurun_out/p21_level3_16384.jsonl"):
    for l in open(f):
        d = json.loads(l)
        print("%-28s value %7.2f  e2e %7.2f  frac %.3f  cpu %s  launches/step %d" % (d["metric"], d["value"], d["e2e"]["value"], d["roofline"]["frac"], ("%.4f" % d["cpu_baseline"]["value"]) if d["cpu_baseline"]["value"] else d["cpu_baseline"]["sample"][:60], d["roofline"]["launches_per_step"]))
PY
