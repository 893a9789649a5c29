#!/bin/bash
set -u
timeout 200 python scripts/bench_dw.py 50 2>&1 | tail -4
timeout 300 python -m pytest tests -m gpu -x -q -k "dwconv or encoder" 2>&1 | tail -2
