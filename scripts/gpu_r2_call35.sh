#!/bin/bash
# 1 GPU: final validation — full suite, smoke, default bench (with extras and CPU baselines)
mkdir -p gpurun_out
echo "== full suite"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/r2c35_full.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r2c35_smoke.txt
echo "== bench default"; timeout 600 python bench.py > gpurun_out/r2c35_bench.json 2> gpurun_out/r2c35_bench.err; grep '^{' gpurun_out/r2c35_bench.json | python scripts/fmt_bench.py
grep '^{' gpurun_out/r2c35_bench.json | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
for w in d.get('other_workloads', []): print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in w.items() if k in ('workload', 'ms_per_step', 'value', 'kernel_ms', 'error')}, 'frac', round(w.get('roofline', {}).get('frac', 0), 3))
print('b1', d.get('operating_points'))
print('roofline', d.get('roofline')); print('e2e', d.get('e2e')); print('clocks', d.get('clocks'))
"
