#!/bin/bash
set -u
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "dwconv" 2>&1 | tail -3
timeout 120 python scripts/bench_dw.py 30 2>&1 | tail -5
