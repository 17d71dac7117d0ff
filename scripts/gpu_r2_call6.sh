#!/bin/bash
# Round-2 GPU call 6 (1 GPU): suite with the fused linear output stage + aligned carve-outs, per-mode throughput, bench
mkdir -p gpurun_out
echo "== full suite"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/r2c6_full.txt
echo "== per-mode throughput"; timeout 600 python scripts/gpu_mode_throughput.py 2>&1 | tail -8 | tee gpurun_out/r2c6_modes.txt
echo "== bench"; timeout 600 python bench.py --no-cpu-baseline --steps 50 --e2e-steps 10 > gpurun_out/r2c6_bench.json 2>gpurun_out/r2c6_bench.err; python scripts/fmt_bench.py < gpurun_out/r2c6_bench.json
echo "== launch list USB"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file gpurun_out/launches_r2_usb.csv python scripts/gpu_mode_throughput.py > /dev/null 2>&1; grep -c . gpurun_out/launches_r2_usb.csv
