#!/bin/bash
# 4 GPUs: why is the step slower than at the old defaults? timeline + option matrix
N=${1:-4}
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
run() {  # name, timeline?, env...
  name=$1; tl=$2; shift; shift
  echo "== $name"
  extra=""; [ "$tl" = "1" ] && extra="--timeline gpurun_out/r2c33_${N}_$name"
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus $N --steps 40 --warmup 3 --e2e-steps 3 --no-weak --no-extras --no-cpu-baseline $extra \
    > gpurun_out/r2c33_${N}_$name.json 2> gpurun_out/r2c33_${N}_$name.err
  grep '^{' gpurun_out/r2c33_${N}_$name.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms/step', round(d['ms_per_step'],4), 'value %.3e' % d['value'], {k: round(v,4) for k,v in d['class_ms_per_step'].items() if v})
" || tail -5 gpurun_out/r2c33_${N}_$name.err
  if [ "$tl" = "1" ]; then python scripts/timeline_print.py gpurun_out/r2c33_${N}_$name 2>/dev/null | sed -n 1,22p; fi
  true
}
run default_tl 1
run kernel 0 KA9Q_B200_MGPU_CE=0
run noprio 0 KA9Q_B200_FFT_PRIO=0
run b2 0 KA9Q_B200_SPEC_BUFFERS=2
run old 0 KA9Q_B200_SPEC_BUFFERS=2 KA9Q_B200_FFT_PRIO=0 KA9Q_B200_MGPU_CE=0
