#!/bin/bash
set -u
OMR_DECODE_WIDE=1 timeout 300 python -m pytest tests -m gpu -x -q -k "greedy or decode or weighted or late" 2>&1 | tail -2
for ph in 0 1 2; do
OMR_DECODE_DBG_PHASE=$ph OMR_DECODE_WIDE=1 timeout 200 python scripts/decode_timing.py 400 2>&1 | tail -3
done
