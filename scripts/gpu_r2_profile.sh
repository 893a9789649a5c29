#!/bin/bash
# round-2 profile: launch list of eager training steps (shares) + `ncu --set full` captures of the dominant kernels
set -u
CMD="python bench.py --steps 1 --warmup 3 --no-cpu --no-decode --no-graph --no-library --modality-drop 0"
timeout 300 $CMD > gpurun_out/plain_r2.log 2> gpurun_out/plain_r2.err &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_r02.csv $CMD > gpurun_out/ncu_r2_list.log 2>&1
echo "launch list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'attn_bwd_tc_kernel' -s 24 -c 4 -o gpurun_out/prof_r2_attn_bwd $CMD > gpurun_out/ncu_r2_a.log 2>&1
echo "attn_bwd capture rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'wgrad_tc_kernel' -s 10 -c 6 -o gpurun_out/prof_r2_wgrad $CMD > gpurun_out/ncu_r2_w.log 2>&1
echo "wgrad capture rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'conv_halo_kernel' -s 20 -c 8 -o gpurun_out/prof_r2_conv_halo $CMD > gpurun_out/ncu_r2_c.log 2>&1
echo "conv capture rc=$?"
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_r02.csv
