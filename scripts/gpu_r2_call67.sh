#!/bin/bash
set -u
timeout 500 python scripts/decode_sweep2.py 400 2>&1 | grep -v Warn | tail -20
timeout 200 python -m pytest tests -m gpu -x -q -k "greedy" 2>&1 | tail -3
