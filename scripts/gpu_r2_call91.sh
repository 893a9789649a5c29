#!/bin/bash
set -u
for ph in 0 1 2; do OMR_DECODE_DBG_PHASE=$ph timeout 200 python scripts/decode_timing.py 400 2>&1 | tail -3 | head -2; done
