#!/bin/bash
set -u
CMD="python bench.py --steps 1 --warmup 3 --no-cpu --no-decode --no-graph --no-library"
$CMD > gpurun_out/plain6.log 2> gpurun_out/plain6.err &&
ncu --set full --clock-control none --import-source on -k regex:'wgrad_tc_kernel' -s 50 -c 6 -o gpurun_out/prof_r1c $CMD > gpurun_out/ncu7.log 2>&1
echo "wgrad capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'conv_halo_kernel' -s 58 -c 5 -o gpurun_out/prof_r1d $CMD > gpurun_out/ncu8.log 2>&1
echo "conv capture rc=$?"
