#!/bin/bash
set -u
OMR_DECODE_WIDE=1 OMR_DECODE_PF_KIND=1 timeout 300 python -m pytest tests -m gpu -x -q -k "greedy or decode or weighted or late" 2>&1 | tail -2
SWEEP_CFGS='[{"OMR_DECODE_WIDE": 1}, {"OMR_DECODE_WIDE": 1, "OMR_DECODE_PF_KIND": 1}, {"OMR_DECODE_WIDE": 1, "OMR_DECODE_PF_KIND": 1, "OMR_DECODE_PF_MASK": "0x7f"}, {"OMR_DECODE_WIDE": 1, "OMR_DECODE_PF_KIND": 1, "OMR_DECODE_PF_MASK": "0x0f"}, {"OMR_DECODE_WIDE": 1, "OMR_DECODE_PF_KIND": 1, "OMR_DECODE_PF_MASK": "0x01"}, {"OMR_DECODE_WIDE": 1, "OMR_DECODE_PF_KIND": 1, "OMR_DECODE_PF_MASK": "0x7f", "OMR_DECODE_NB": 3}, {"OMR_DECODE_WIDE": 1, "OMR_DECODE_PF_KIND": 1, "OMR_DECODE_PF_MASK": "0x60"}]' timeout 300 python scripts/decode_sweep2.py 400 2>&1 | grep -v Warn | tail -8
OMR_DECODE_PF_KIND=1 OMR_DECODE_WIDE=1 OMR_DECODE_DBG_PHASE=2 timeout 200 python scripts/decode_timing.py 400 2>&1 | tail -3
