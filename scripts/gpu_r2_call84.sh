#!/bin/bash
set -u
SWEEP_CFGS='[{}, {"OMR_DECODE_PF_MASK": "0x7f"}, {"OMR_DECODE_PF_MASK": "0x30"}, {"OMR_DECODE_PF_MASK": "0x78"}, {"OMR_DECODE_PF_MASK": "0x7c"}, {"OMR_DECODE_PF_MASK": "0x60"}, {"OMR_DECODE_PF_MASK": "0x20"}, {"OMR_DECODE_PF_SELF": 0}, {"OMR_DECODE_PF_SELF": 512}, {"OMR_DECODE_PF_CROSS": 1600}, {}]' timeout 400 python scripts/decode_sweep2.py 1268 2>&1 | grep -v Warn | tail -11
