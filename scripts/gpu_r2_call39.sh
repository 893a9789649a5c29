#!/bin/bash
set -u
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q -x -k "proj_ce" > gpurun_out/t_pce.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/t_pce.log | cut -c1-220
