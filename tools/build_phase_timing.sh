#!/bin/sh
# Developer tool: build t41_sdr_b200/libt41rx_ptiming.so = libt41rx with -DT41RX_PHASE_TIMING (cycles per phase of CTA 0 of
# the bit-exact kernel, slots 0..31, and of the rows kernel, slots 32..63; read with t41rx_debug_phase_cycles).
cd "$(dirname "$0")/../t41_sdr_b200/csrc" || exit 1
make -s all || exit 1
nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo --fmad=false -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math,-O2 \
     -DT41RX_PHASE_TIMING "$@" -c rx_api.cu -o /tmp/rx_api_ptiming.o || exit 1
nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo --fmad=false -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math,-O2 \
     -DT41RX_PHASE_TIMING "$@" -c rx_rows.cu -o /tmp/rx_rows_ptiming.o || exit 1
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libt41rx_ptiming.so /tmp/rx_api_ptiming.o /tmp/rx_rows_ptiming.o rx_fast.o rx_design.o rx_host.o rx_wav.o rx_multi.o -ldl || exit 1
echo built libt41rx_ptiming.so
