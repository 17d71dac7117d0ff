#!/bin/bash
# Sweep the big-FFT launch knobs on one box: prints the fft class time per step for each setting.
run() { python bench.py --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); print('$1', 'ms/step %.4f' % d['ms_per_step'], {k: round(x, 4) for k, x in d['class_ms_per_step'].items() if x})"; }
for c in 2 3 4 5 6 8; do KA9Q_B200_FFT_CTAS=$c run "async ctas=$c"; done
KA9Q_B200_FFT_ASYNC=0 run "sync"
