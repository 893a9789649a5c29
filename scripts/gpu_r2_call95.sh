#!/bin/bash
set -u
SWEEP_CFGS='[{}, {}, {}]' timeout 400 python scripts/decode_sweep2.py 1268 2>&1 | grep -v Warn | tail -3
