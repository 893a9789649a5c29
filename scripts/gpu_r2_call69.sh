#!/bin/bash
set -u
timeout 300 python -m pytest tests -m gpu -x -q -k "greedy or decode or weighted or late" 2>&1 | tail -5
SWEEP_CFGS='[{}, {"OMR_DECODE_PF_MASK": "0x70"}, {"OMR_DECODE_PF_MASK": "0x7f"}, {"OMR_DECODE_NB": 3, "OMR_DECODE_PF_MASK": "0x7f"}]' timeout 300 python scripts/decode_sweep2.py 400 2>&1 | grep -v Warn | tail -6
timeout 200 python scripts/decode_timing.py 400 2>&1 | tail -4
