#!/bin/bash
set -u
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_ops.py -m gpu -q -x -k "attention or attn" > gpurun_out/t_attn.log 2>&1; echo "pytest attn rc=$?"; tail -4 gpurun_out/t_attn.log
for d in 256; do echo "OMR_ATTN_DEBUG=$d"; OMR_ATTN_DEBUG=$d timeout 200 python scripts/attn_stamps.py 2>&1 | tail -8 | cut -c1-200; done
for d in 0 2 62; do echo "OMR_ATTN_DEBUG=$d"; OMR_ATTN_DEBUG=$d timeout 200 python scripts/bench_attn.py 20 2>&1 | grep "self\|cross"; done
