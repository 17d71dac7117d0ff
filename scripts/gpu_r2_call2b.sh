#!/bin/bash
# Round-2 GPU call 2b (2 GPUs): multi-GPU parity (P2P / NCCL / broadcast vs one GPU) and the sharded bench at N=2
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
nvidia-smi topo -m > gpurun_out/r2c2b_topo.txt 2>&1
echo "== pytest multi-GPU"; timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -30 | tee gpurun_out/r2c2b_pytest.txt
for tr in p2p nccl; do
  echo "== bench --gpus 2 sharded $tr"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 50 --warmup 3 --e2e-steps 20 --transport $tr > gpurun_out/r2c2b_bench_$tr.json 2> gpurun_out/r2c2b_bench_$tr.err
  tail -c 2500 gpurun_out/r2c2b_bench_$tr.json; tail -5 gpurun_out/r2c2b_bench_$tr.err
done
