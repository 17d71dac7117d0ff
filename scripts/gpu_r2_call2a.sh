#!/bin/bash
# Round-2 GPU call 2a (1 GPU): the new BASELINE-geometry parity tests, A/B of the shared-memory stage-2 twiddle build
mkdir -p gpurun_out
echo "== pytest new parity tests"; timeout 1200 python -m pytest tests/test_gpu_parity_configs.py -m gpu -q -s 2>&1 | tail -30 | tee gpurun_out/r2c2a_pytest_cfg.txt
echo "== pytest full suite (current default build)"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 | tee gpurun_out/r2c2a_pytest.txt
echo "== A/B variants"
export AB_BENCH_ARGS="--steps 50 --e2e-steps 10 --warmup 3"
timeout 900 scripts/ab_variants.sh tw s8 s7 s7e s8e s7c60 2>&1 | tee gpurun_out/r2c2a_ab.txt
