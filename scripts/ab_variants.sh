#!/bin/bash
# Same-box A/B of kernel variants: variants/<name>.so are prebuilt copies of libka9q_b200.so (see KA9Q_B200_NVCC_EXTRA
# in ka9q_sdr_b200/build.py). Usage (on the GPU box): scripts/ab_variants.sh A E F ...   -> one line per variant.
cp ka9q_sdr_b200/libka9q_b200.so /tmp/lib_keep.so
for v in "$@"; do
  cp variants/$v.so ka9q_sdr_b200/libka9q_b200.so
  python bench.py --no-cpu-baseline --no-extras ${AB_BENCH_ARGS} 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
print('$v', 'ms/step %.4f' % d['ms_per_step'], 'e2e %.3f' % d['e2e']['ms_per_step'], {k: round(x, 4) for k, x in d['class_ms_per_step'].items() if x})
"
done
cp /tmp/lib_keep.so ka9q_sdr_b200/libka9q_b200.so
