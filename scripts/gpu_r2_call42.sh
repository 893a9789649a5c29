#!/bin/bash
set -u
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "dropout or instnorm or masks" 2>&1 | tail -3
for v in 0 1 2 3; do echo "OMR_IN_VARIANT=$v"; OMR_IN_VARIANT=$v timeout 200 python scripts/bench_stream.py 10 2>&1 | tail -3; done
