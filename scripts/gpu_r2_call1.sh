#!/bin/bash
# Round-2 GPU call 1: full GPU test suite (incl. the new BASELINE-geometry parity tests), same-box A/B of fm_kernel
# variants, ncu full capture + launch list of the default build. Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2c1_smi.txt 2>&1
echo "== pytest -m gpu (full)"; timeout 900 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -25 | tee gpurun_out/r2c1_pytest.txt
echo "== A/B variants"
export AB_BENCH_ARGS="--steps 50 --e2e-steps 10 --warmup 3"
timeout 900 scripts/ab_variants.sh base tw tw_c7 tma1 tma2 2>&1 | tee gpurun_out/r2c1_ab.txt
echo "== parity of the TMA variants (golden chains, cfg3, cfg5 full-size properties)"
cp ka9q_sdr_b200/libka9q_b200.so /tmp/lib_keep.so
for v in tma1 tma2; do
  cp variants/$v.so ka9q_sdr_b200/libka9q_b200.so
  echo "-- $v"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "golden or cfg3 or cfg5_full or long_run or wraparound" 2>&1 | tail -4 | tee -a gpurun_out/r2c1_tma_parity.txt
done
cp /tmp/lib_keep.so ka9q_sdr_b200/libka9q_b200.so
echo "== bench default"; timeout 600 python bench.py > gpurun_out/r2c1_bench.json 2> gpurun_out/r2c1_bench.err; tail -c 3000 gpurun_out/r2c1_bench.json
echo "== ncu full fm_kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fm_kernel -s 8 -c 1 -f -o gpurun_out/prof_fm_r2a \
  python bench.py --no-cpu-baseline --steps 3 --warmup 3 --e2e-steps 2 > gpurun_out/r2c1_ncu_fm.log 2>&1; tail -3 gpurun_out/r2c1_ncu_fm.log
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 200 --csv --log-file gpurun_out/launches_r2a.csv \
  python bench.py --no-cpu-baseline --steps 3 --warmup 3 --e2e-steps 2 > gpurun_out/r2c1_ncu_ll.log 2>&1; tail -2 gpurun_out/r2c1_ncu_ll.log
ls -la gpurun_out | tail -12
