#!/bin/bash
# attention kernels v2: tests that touch attention + microbenchmark
set -u
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_ops.py -m gpu -q -x -k "attention or attn" > gpurun_out/t_attn.log 2>&1; echo "pytest attn rc=$?"; tail -15 gpurun_out/t_attn.log
timeout 200 python scripts/bench_attn.py 20 > gpurun_out/attn_v2.log 2>&1; echo "attn rc=$?"; cat gpurun_out/attn_v2.log
