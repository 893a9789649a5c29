#!/bin/bash
set -u
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -q -x -k "attention" 2>&1 | tail -4
