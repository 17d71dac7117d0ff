#!/bin/bash
# AddressSanitizer + UBSan pass over the host C sources (rtp_glue.c, rx_host.c, osc_host.c): builds an instrumented copy of
# the library out of tree (the CUDA objects are reused as built) and runs the CPU tests that drive those sources.
# Leak checking is off (the Python interpreter itself is not leak-clean under ASan); overflows, use-after-free and UB are on.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
W=/tmp/k9_asan_build
rm -rf $W; mkdir -p $W
python -m ka9q_sdr_b200.build > /dev/null
cp -r "$ROOT/ka9q_sdr_b200" "$ROOT/include" "$ROOT/oracle" "$ROOT/tests" $W/
cd $W/ka9q_sdr_b200
for f in osc_host rtp_glue rx_host; do
  gcc -O1 -g -std=gnu11 -fPIC -fsanitize=address,undefined -fno-omit-frame-pointer -c csrc/$f.c -o build/$f.c.o
done
nvcc -shared -o libka9q_b200.so build/*.o -gencode arch=compute_100a,code=sm_100a -Xlinker -Bsymbolic -lpthread -ldl -lm \
  -Xlinker -lasan -Xlinker -lubsan
cd $W
LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 UBSAN_OPTIONS=print_stacktrace=1 \
  python -m pytest tests/test_rtp_glue.py tests/test_rx_host.py tests/test_abi.py -x -q
rm -rf $W
