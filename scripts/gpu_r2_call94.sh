#!/bin/bash
set -u
SWEEP_CFGS='[{}, {"OMR_DECODE_NB": 3}, {}, {"OMR_DECODE_NB": 3}]' timeout 400 python scripts/decode_sweep2.py 1268 2>&1 | grep -v Warn | tail -4
