#!/bin/bash
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
for n in 148 296 592 1184; do
  echo "== scatter ctas $n"
  KA9Q_B200_SCATTER_CTAS=$n timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus 2 --steps 50 --warmup 3 --e2e-steps 5 --no-weak 2>/dev/null | grep '^{' | python scripts/fmt_bench.py | tee -a gpurun_out/r2c13_scatter.txt
done
