#!/bin/bash
set -u
timeout 400 python -m pytest tests -m gpu -x -q -k "greedy or decode or weighted or late" 2>&1 | tail -2
SWEEP_CFGS='[{}, {}, {}]' timeout 400 python scripts/decode_sweep2.py 1268 2>&1 | grep -v Warn | tail -3
