#!/bin/bash
set -u
for d in 0 1 2 4 3 7; do
  echo "== OMR_CONV_DEBUG=$d (fwd only, few shapes)"
  BENCH_CONV_FEW=1 OMR_CONV_DEBUG=$d timeout 120 python scripts/bench_conv.py fwd 2>&1 | tail -8
done
for d in 8 10; do echo "== stamps OMR_CONV_DEBUG=$d"; OMR_CONV_DEBUG=$d python scripts/conv_stamps.py 2>&1 | grep -A14 "Cin 16" | head -16; done
