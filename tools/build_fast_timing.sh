#!/bin/sh
# Developer tool: build t41_sdr_b200/libt41rx_timing.so = libt41rx with -DT41RX_FAST_TIMING (tools/fast_timing.py).
cd "$(dirname "$0")/../t41_sdr_b200/csrc" || exit 1
make -s all || exit 1
nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math,-O2 \
     -DT41RX_FAST_TIMING "$@" -c rx_fast.cu -o /tmp/rx_fast_timing.o || exit 1
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libt41rx_timing.so rx_api.o rx_rows.o /tmp/rx_fast_timing.o rx_design.o rx_host.o rx_wav.o rx_multi.o -ldl || exit 1
echo built libt41rx_timing.so
