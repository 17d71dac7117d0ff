#!/bin/bash
mkdir -p gpurun_out
echo "== PL test"; timeout 600 python -m pytest tests/test_gpu_parity_configs.py -m gpu -q -k "pl_tone" 2>&1 | tail -12 | tee gpurun_out/r2c17_pl.txt
echo "== launch list USB"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 60 --csv --log-file gpurun_out/launches_r2_usb.csv python scripts/gpu_usb_only.py USB > /dev/null 2>&1
echo "== launch list AM"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 60 --csv --log-file gpurun_out/launches_r2_am.csv python scripts/gpu_usb_only.py AM > /dev/null 2>&1
python - <<'PY'
import csv, collections
for f in ("usb", "am"):
    rows = list(csv.reader(open(f"gpurun_out/launches_r2_{f}.csv")))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hi]; kn = h.index("Kernel Name"); mv = h.index("Metric Value")
    acc = collections.defaultdict(list)
    for r in rows[hi + 2:]:
        if len(r) > mv:
            try: acc[r[kn][:50]].append(float(r[mv].replace(",", "")))
            except Exception: pass
    print(f, {k: (len(v), round(sum(v) / len(v) / 1000, 1)) for k, v in acc.items()})
PY
