#!/bin/bash
# Round-2 GPU call 2c (1 GPU): A/B of the shared-memory stage-2 twiddle builds, FFT blocks-per-launch sweep, n0 parity test
mkdir -p gpurun_out
echo "== n0 parity test"; timeout 600 python -m pytest tests/test_gpu_parity_configs.py -m gpu -q -s -k n0 2>&1 | tail -15 | tee gpurun_out/r2c2c_pytest_n0.txt
echo "== A/B variants"
export AB_BENCH_ARGS="--steps 50 --e2e-steps 10 --warmup 3"
timeout 900 scripts/ab_variants.sh s8 s7 s7e s8e s7c60 2>&1 | tee gpurun_out/r2c2c_ab.txt
echo "== FFT blocks per launch (default build)"
for n in 1 2 4; do
  KA9Q_B200_FFT_BLOCKS_PER_LAUNCH=$n python bench.py --no-cpu-baseline --steps 50 --e2e-steps 10 --warmup 3 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
print('fft_blocks_per_launch=$n', 'ms/step %.4f' % d['ms_per_step'], {k: round(x, 4) for k, x in d['class_ms_per_step'].items() if x})
" | tee -a gpurun_out/r2c2c_fftbpl.txt
done
