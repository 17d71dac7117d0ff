#!/bin/bash
# 4 GPUs: exchange on its own stream. Parity of three P2P modes (2 of the GPUs), bench; the copy-kernel variant only if
# the default is not already good
N=${1:-4}
mkdir -p gpurun_out
export CUDA_DEVICE_MAX_CONNECTIONS=32
echo "== pytest multi-GPU (p2p, p2p-stream, p2p-kernel)"
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "p2p] or p2p-stream or p2p-kernel" 2>&1 | tail -3
run() {  # name, env...
  name=$1; shift
  echo "== $name"
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus $N --steps 40 --warmup 3 --e2e-steps 3 --no-weak --no-extras --no-cpu-baseline --timeline gpurun_out/r2c34_${N}_$name \
    > gpurun_out/r2c34_${N}_$name.json 2> gpurun_out/r2c34_${N}_$name.err
  grep '^{' gpurun_out/r2c34_${N}_$name.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms/step', round(d['ms_per_step'],4), 'value %.3e' % d['value'], {k: round(v,4) for k,v in d['class_ms_per_step'].items() if v})
open('gpurun_out/r2c34_last_ms','w').write(str(d['ms_per_step']))
" || tail -5 gpurun_out/r2c34_${N}_$name.err
  python scripts/timeline_print.py gpurun_out/r2c34_${N}_$name 2>/dev/null | sed -n 1,18p
  true
}
run default
if python -c "import sys; sys.exit(0 if float(open('gpurun_out/r2c34_last_ms').read()) > 0.150 else 1)"; then
  run kernel KA9Q_B200_MGPU_CE=0
fi
