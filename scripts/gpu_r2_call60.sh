#!/bin/bash
set -u
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_ops.py -m gpu -q -x -k "attention or attn" 2>&1 | tail -4
timeout 200 python scripts/bench_attn.py 20 2>&1 | tail -4
OMR_ATTN_DEBUG=512 timeout 100 python scripts/attn_stamps_fwd.py drop 2>&1 | tail -6 | cut -c1-150
