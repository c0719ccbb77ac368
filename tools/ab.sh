#!/bin/sh
# Developer tool: A/B two builds of libt41rx on the SAME GPU box (box-to-box variance is ~2 %).
#   tools/ab.sh save NAME      copy the current build to t41_sdr_b200/libt41rx_NAME.so
#   tools/ab.sh run A B [bench args]   (on the GPU box) alternate the two builds, 3 rounds each
cd "$(dirname "$0")/.." || exit 1
case "$1" in
  save) make -s -C t41_sdr_b200/csrc all && cp t41_sdr_b200/libt41rx.so "t41_sdr_b200/libt41rx_$2.so" && echo "saved libt41rx_$2.so" ;;
  run)
    A=$2; B=$3; shift 3
    for r in 1 2 3; do
      for v in "$A" "$B"; do
        T41RX_LIB="$PWD/t41_sdr_b200/libt41rx_$v.so" python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline "$@" 2>&1 | tail -1 |
          python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['value']), round(d['roofline']['avg_launch_ms'],4))"
      done
    done ;;
esac
