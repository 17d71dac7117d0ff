#!/bin/bash
# Round-2 GPU call 5 (N GPUs): PCIe probe with 1..N GPUs copying at once, then the sharded bench at N
N=${1:-8}
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
nvidia-smi topo -m > gpurun_out/r2c5_topo_$N.txt 2>&1
echo "== pcie probe"; timeout 300 python scripts/gpu_pcie_probe_multi.py --gpus $N 2>&1 | tail -2 | tee gpurun_out/r2c5_pcie_$N.json
echo "== bench --gpus $N sharded p2p"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 \
  bench.py --gpus $N --steps 50 --warmup 3 --e2e-steps 20 > gpurun_out/r2c5_bench_$N.json 2> gpurun_out/r2c5_bench_$N.err
tail -c 2600 gpurun_out/r2c5_bench_$N.json; tail -3 gpurun_out/r2c5_bench_$N.err
