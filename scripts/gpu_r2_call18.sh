#!/bin/bash
# validation of the committed state (wgrad scratch path through the tests) + attention microbenchmark baseline
set -u
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/t_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/t_gpu.log
timeout 200 python scripts/bench_attn.py 20 > gpurun_out/attn_base.log 2>&1; echo "attn rc=$?"; cat gpurun_out/attn_base.log
