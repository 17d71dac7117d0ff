#!/bin/bash
# 1 GPU: final validation — full suite, smoke, default bench (with extras and CPU baselines), reference arm, launch list
mkdir -p gpurun_out
echo "== full suite"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/r2c15_full.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r2c15_smoke.txt
echo "== bench default"; /usr/bin/time -v timeout 900 python bench.py > gpurun_out/r2c15_bench.json 2> gpurun_out/r2c15_bench.err; grep -E "Elapsed|Maximum resident" gpurun_out/r2c15_bench.err; grep '^{' gpurun_out/r2c15_bench.json | python scripts/fmt_bench.py
grep '^{' gpurun_out/r2c15_bench.json | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
for w in d.get('other_workloads', []): print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in w.items() if k in ('workload', 'ms_per_step', 'value', 'kernel_ms', 'error')}, 'frac', round(w.get('roofline', {}).get('frac', 0), 3))
print('b1', d.get('operating_points'))
"
echo "== launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file gpurun_out/launches_r2b.csv python bench.py --no-cpu-baseline --no-extras --steps 3 --warmup 3 --e2e-steps 2 > /dev/null 2>&1; wc -l gpurun_out/launches_r2b.csv
