#!/bin/bash
set -u
for d in 0 0 1 2 4 7; do
  echo "== OMR_CONV_DEBUG=$d"
  BENCH_CONV_FEW=1 OMR_CONV_DEBUG=$d timeout 120 python scripts/bench_conv.py fwd 2>&1 | tail -7 | head -6
done
