#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke15.log 2>&1; tail -5 gpurun_out/smoke15.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu15.log 2>&1
tail -4 gpurun_out/pytest_gpu15.log
python bench.py > gpurun_out/bench15_dgemm16384.json 2> gpurun_out/bench15.err
for w in dgemm_rankk zgemm4096 sgemm8192 cgemm4096; do
  timeout 600 python bench.py --steps 5 --warmup 3 --workload $w > gpurun_out/bench15_$w.json 2>> gpurun_out/bench15.err
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench15_ref.json 2>> gpurun_out/bench15.err
tail -3 gpurun_out/bench15.err
for f in gpurun_out/bench15_*.json; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().split('\n')[-1])
r=d.get('roofline') or {}
print(sys.argv[1].split('bench15_')[1], 'value %.2f'%d['value'], 'ms %.3f'%d['ms_per_step'], 'frac', r.get('frac'), 'e2e', (d.get('e2e') or {}).get('value'), 'cpu', (d.get('cpu_baseline') or {}).get('value'), 'launches', d.get('gpu_launches'), d.get('kernel'))
PY
done
