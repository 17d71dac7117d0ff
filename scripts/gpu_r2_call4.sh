#!/bin/bash
# Round-2 GPU call 4 (1 GPU): full suite on the shared-response build, A/B, bench, ncu of fm_kernel + launch list
mkdir -p gpurun_out
echo "== full suite"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/r2c4_full.txt
echo "== A/B variants"
export AB_BENCH_ARGS="--steps 50 --e2e-steps 10 --warmup 3"
timeout 900 scripts/ab_variants.sh d8 d8e d7 2>&1 | tee gpurun_out/r2c4_ab.txt
echo "== bench default"; timeout 600 python bench.py > gpurun_out/r2c4_bench.json 2> gpurun_out/r2c4_bench.err; tail -c 1500 gpurun_out/r2c4_bench.json
echo "== ncu full fm_kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fm_kernel -s 8 -c 1 -f -o gpurun_out/prof_fm_r2b \
  python bench.py --no-cpu-baseline --steps 3 --warmup 3 --e2e-steps 2 > gpurun_out/r2c4_ncu_fm.log 2>&1; tail -2 gpurun_out/r2c4_ncu_fm.log
