#!/bin/bash
set -u
python scripts/wgrad_one.py 16 128 128 128 1 1 > gpurun_out/wg1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:wgrad -s 2 -c 1 -o gpurun_out/prof_wgrad_c128 python scripts/wgrad_one.py 16 128 128 128 1 1 > gpurun_out/ncu_wg1.log 2>&1
tail -2 gpurun_out/ncu_wg1.log
python scripts/wgrad_one.py 128 1024 32 32 1 1 > gpurun_out/wg2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:wgrad -s 2 -c 1 -o gpurun_out/prof_wgrad_c32 python scripts/wgrad_one.py 128 1024 32 32 1 1 > gpurun_out/ncu_wg2.log 2>&1
tail -2 gpurun_out/ncu_wg2.log
