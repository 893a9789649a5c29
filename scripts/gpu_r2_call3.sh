#!/bin/bash
set -u
for d in 8 15; do echo "== stamps OMR_CONV_DEBUG=$d"; OMR_CONV_DEBUG=$d python scripts/conv_stamps.py 2>&1 | tail -60; done
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/t_gpu.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/t_gpu.log
