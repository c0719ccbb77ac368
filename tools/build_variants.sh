#!/bin/sh
# Developer tool: build tuning variants of libt41rx (receivers per CTA G, serial-lane mapping W, phase timing).
#   tools/build_variants.sh "G W [extra nvcc defines]" ...   ->   t41_sdr_b200/libt41rx_g<G>w<W><tag>.so
cd "$(dirname "$0")/../t41_sdr_b200/csrc" || exit 1
make -s all || exit 1
for spec in "$@"; do
  set -- $spec
  G=$1; W=$2; shift 2
  tag=""; [ -n "$*" ] && tag="_$(echo "$*" | tr -dc 'A-Za-z0-9')"
  nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo --fmad=false \
       -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math,-O2 -DT41RX_G=$G -DT41RX_SERIAL_WPS=$W "$@" \
       -c rx_api.cu -o /tmp/rx_api_variant.o || exit 1
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libt41rx_g${G}w${W}${tag}.so /tmp/rx_api_variant.o rx_design.o rx_host.o || exit 1
  echo built libt41rx_g${G}w${W}${tag}.so
done
