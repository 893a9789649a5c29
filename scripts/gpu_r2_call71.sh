#!/bin/bash
set -u
OMR_DECODE_WIDE=1 timeout 300 python -m pytest tests -m gpu -x -q -k "greedy or decode or weighted or late" 2>&1 | tail -2
SWEEP_CFGS='[{}, {"OMR_DECODE_WIDE": 1}, {"OMR_DECODE_WIDE": 1, "OMR_DECODE_NB": 3}, {"OMR_DECODE_WIDE": 1, "OMR_DECODE_NB": 3, "OMR_DECODE_STAGGER_NS": 2000}, {"OMR_DECODE_WIDE": 1, "OMR_DECODE_PF_CROSS": 0}]' timeout 300 python scripts/decode_sweep2.py 400 2>&1 | grep -v Warn | tail -6
OMR_DECODE_WIDE=1 DECODE_BATCH=8 timeout 200 python scripts/decode_timing.py 300 2>&1 | tail -2
OMR_DECODE_WIDE=1 OMR_DECODE_NB=3 DECODE_BATCH=8 timeout 200 python scripts/decode_timing.py 300 2>&1 | tail -2
