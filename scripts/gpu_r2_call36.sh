#!/bin/bash
# 8 GPUs: one strong-scaling line at the final defaults
N=8
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
  bench.py --gpus $N --steps 60 --warmup 3 --e2e-steps 3 --no-weak --no-extras --no-cpu-baseline \
  > gpurun_out/r2c36_bench_8.json 2> gpurun_out/r2c36_bench_8.err
grep '^{' gpurun_out/r2c36_bench_8.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms/step', round(d['ms_per_step'],4), 'value %.3e' % d['value'], {k: round(v,4) for k,v in d['class_ms_per_step'].items() if v})
" || tail -5 gpurun_out/r2c36_bench_8.err
