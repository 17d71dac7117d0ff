#!/bin/bash
# N GPUs ($1): overlap experiments on the sharded strong-scaling step. For each setting: bench (short) + region timeline.
N=${1:-2}
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
run() {  # name, env...
  name=$1; shift
  echo "== $name"
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus $N --steps 60 --warmup 3 --e2e-steps 3 --no-weak --no-extras --no-cpu-baseline --timeline gpurun_out/r2c27_${N}_$name \
    > gpurun_out/r2c27_${N}_$name.json 2> gpurun_out/r2c27_${N}_$name.err
  grep '^{' gpurun_out/r2c27_${N}_$name.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms/step', round(d['ms_per_step'],4), 'value %.3e' % d['value'], {k: round(v,4) for k,v in d['class_ms_per_step'].items() if v})
"
  python scripts/timeline_print.py gpurun_out/r2c27_${N}_$name | head -45
}
run base KA9Q_B200_FFT_PRIO=0
run prio KA9Q_B200_FFT_PRIO=1
run prio_ce KA9Q_B200_FFT_PRIO=1 KA9Q_B200_MGPU_CE=1
