#!/bin/bash
mkdir -p gpurun_out
echo "== PL test"; timeout 600 python -m pytest tests/test_gpu_parity_configs.py -m gpu -q -k "pl_tone" 2>&1 | tail -25 | tee gpurun_out/r2c16_pl.txt
echo "== suite"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
echo "== bench default"; timeout 900 python bench.py > gpurun_out/r2c15_bench.json 2> gpurun_out/r2c15_bench.err; tail -3 gpurun_out/r2c15_bench.err; grep '^{' gpurun_out/r2c15_bench.json | python scripts/fmt_bench.py
grep '^{' gpurun_out/r2c15_bench.json | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
for w in d.get('other_workloads', []): print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in w.items() if k in ('workload', 'ms_per_step', 'value', 'kernel_ms', 'error')}, 'frac', round(w.get('roofline', {}).get('frac', 0), 3))
print('b1', d.get('operating_points'))
"
