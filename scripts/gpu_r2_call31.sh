#!/bin/bash
# N GPUs ($1): multi-GPU parity (when $2 = test) and the full default bench line (weak line included)
N=${1:-2}
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
if [ "$2" = "test" ]; then
  echo "== pytest multi-GPU"; timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -4
fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 \
  bench.py --gpus $N --steps 60 --warmup 3 --e2e-steps 20 --no-extras > gpurun_out/r2c31_bench_$N.json 2> gpurun_out/r2c31_bench_$N.err
grep '^{' gpurun_out/r2c31_bench_$N.json | python scripts/fmt_bench.py
