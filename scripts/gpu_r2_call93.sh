#!/bin/bash
set -u
timeout 200 python scripts/decode_timing.py 400 2>&1 | tail -2
SWEEP_CFGS='[{}, {"OMR_DECODE_PF_MASK": "0x7f"}, {"OMR_DECODE_PF_MASK": "0x20"}, {"OMR_DECODE_PF_CROSS": 0}, {"OMR_DECODE_PF_MASK": "0x78"}, {}]' timeout 400 python scripts/decode_sweep2.py 1268 2>&1 | grep -v Warn | tail -6
