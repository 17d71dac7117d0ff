#!/bin/bash
# 8 GPUs: copy-engine vs copy-kernel exchange, 8 vs 16 blocks per step
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
echo "== multi-GPU parity with the copy-engine exchange"; KA9Q_B200_MGPU_CE=1 timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "p2p" 2>&1 | tail -3 | tee gpurun_out/r2c14_pytest_ce.txt
run() { # name, env, extra args
  echo "== $1"
  env $2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 \
    bench.py --gpus 8 --steps 50 --warmup 3 --e2e-steps 10 --no-weak $3 2>/dev/null | grep '^{' > gpurun_out/r2c14_$1.json
  python scripts/fmt_bench.py < gpurun_out/r2c14_$1.json | tee -a gpurun_out/r2c14_summary.txt
}
run kernel_b8 "KA9Q_B200_MGPU_CE=0" ""
run ce_b8 "KA9Q_B200_MGPU_CE=1" ""
run kernel_b16 "KA9Q_B200_MGPU_CE=0" "--blocks 16"
run ce_b16 "KA9Q_B200_MGPU_CE=1" "--blocks 16"
