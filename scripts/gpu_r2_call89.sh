#!/bin/bash
set -u
timeout 600 python -m pytest tests/test_gpu_decode_persistent.py -m gpu -x -q 2>&1 | tail -12
