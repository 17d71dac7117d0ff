#!/bin/bash
# Round-2 GPU call 3 (1 GPU): n0 / PLL / front-end parity, then the whole suite
mkdir -p gpurun_out
echo "== new tests"; timeout 900 python -m pytest tests/test_gpu_parity_configs.py tests/test_gpu_frontend.py -m gpu -q -s -k "n0 or coherent or frontend" 2>&1 | tail -40 | tee gpurun_out/r2c3_new.txt
echo "== full suite"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/r2c3_full.txt
