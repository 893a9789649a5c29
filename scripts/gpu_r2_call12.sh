#!/bin/bash
set -u
export OMR_CONV_STAGE=0
python scripts/conv_one.py 16 16 > gpurun_out/conv_one.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:conv_halo -s 2 -c 1 -o gpurun_out/prof_conv16 python scripts/conv_one.py 16 16 > gpurun_out/ncu_conv16.log 2>&1
tail -3 gpurun_out/ncu_conv16.log
