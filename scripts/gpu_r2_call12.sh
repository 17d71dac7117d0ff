#!/bin/bash
# 2 GPUs: multi-GPU parity after the scatter rewrite, bench at N=2
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
echo "== pytest multi-GPU"; timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -6 | tee gpurun_out/r2c12_pytest.txt
echo "== bench --gpus 2"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
  bench.py --gpus 2 --steps 50 --warmup 3 --e2e-steps 20 > gpurun_out/r2c12_bench_2.json 2> gpurun_out/r2c12_bench_2.err
grep '^{' gpurun_out/r2c12_bench_2.json | python scripts/fmt_bench.py
