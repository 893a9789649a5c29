#!/bin/bash
set -u
timeout 400 python -m pytest tests -m gpu -x -q -k "greedy or decode or weighted or late" 2>&1 | tail -2
SWEEP_CFGS='[{}, {}]' timeout 300 python scripts/decode_sweep2.py 1268 2>&1 | grep -v Warn | tail -2
SWEEP_CFGS='[{}]' timeout 300 python scripts/decode_sweep2.py 400 2>&1 | grep -v Warn | tail -1
OMR_DECODE_DBG_PHASE=2 timeout 200 python scripts/decode_timing.py 400 2>&1 | tail -3
