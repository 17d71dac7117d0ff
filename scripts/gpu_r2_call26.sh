#!/bin/bash
# block-split FM form: bit-identity test, the FM-touching parity tests, then the per-shape timing probe
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_configs.py -m gpu -q -x -k "block_split or golden or long_run or squelch or cfg3 or batching or overlap_geom or off_grid or pl_tone or cfg4" 2>&1 | tail -15
timeout 600 python scripts/gpu_fm_split_probe.py 2>&1 | grep -v '^NCCL' | tee gpurun_out/r2c26_split.txt
