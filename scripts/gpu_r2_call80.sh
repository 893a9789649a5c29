#!/bin/bash
set -u
OMR_DECODE_TAIL=1 timeout 400 python -m pytest tests -m gpu -x -q -k "greedy or decode or weighted or late" 2>&1 | tail -2
SWEEP_CFGS='[{}, {"OMR_DECODE_TAIL": 1}, {}, {"OMR_DECODE_TAIL": 1}]' timeout 300 python scripts/decode_sweep2.py 1268 2>&1 | grep -v Warn | tail -4
OMR_DECODE_TAIL=1 timeout 200 python scripts/decode_timing.py 1268 2>&1 | tail -2
