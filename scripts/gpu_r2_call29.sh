#!/bin/bash
# N GPUs ($1): side-stream flag kernels + FFT priority (defaults now) vs. options; parity first when $2 = test
N=${1:-2}
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
if [ "$2" = "test" ]; then
  echo "== pytest multi-GPU"; timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -4
fi
run() {  # name, timeline?, env...
  name=$1; tl=$2; shift; shift
  echo "== $name"
  extra=""; [ "$tl" = "1" ] && extra="--timeline gpurun_out/r2c29_${N}_$name"
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus $N --steps 60 --warmup 3 --e2e-steps 3 --no-weak --no-extras --no-cpu-baseline $extra \
    > gpurun_out/r2c29_${N}_$name.json 2> gpurun_out/r2c29_${N}_$name.err
  grep '^{' gpurun_out/r2c29_${N}_$name.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms/step', round(d['ms_per_step'],4), 'value %.3e' % d['value'], {k: round(v,4) for k,v in d['class_ms_per_step'].items() if v})
" || tail -5 gpurun_out/r2c29_${N}_$name.err
  [ "$tl" = "1" ] && python scripts/timeline_print.py gpurun_out/r2c29_${N}_$name 2>/dev/null | sed -n 1,22p
}
run default 0
run default_tl 1
run ce 0 KA9Q_B200_MGPU_CE=1
run b2 0 KA9Q_B200_SPEC_BUFFERS=2
run noprio 0 KA9Q_B200_FFT_PRIO=0
