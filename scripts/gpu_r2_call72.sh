#!/bin/bash
set -u
OMR_DECODE_WIDE=1 timeout 300 python -m pytest tests -m gpu -x -q -k "greedy or decode or weighted or late" 2>&1 | tail -12
SWEEP_CFGS='[{}, {"OMR_DECODE_WIDE": 1}]' timeout 300 python scripts/decode_sweep2.py 400 2>&1 | grep -v Warn | tail -3
OMR_DECODE_WIDE=1 timeout 200 python scripts/decode_timing.py 400 2>&1 | tail -3
