#!/bin/bash
# end-of-round evidence: launch list of eager training steps (shares) + `ncu --set full` captures of the rewritten kernels
set -u
CMD="python bench.py --steps 1 --warmup 3 --no-cpu --no-decode --no-graph --no-library --modality-drop 0"
timeout 300 $CMD > gpurun_out/plain_final.log 2> gpurun_out/plain_final.err &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 5000 --csv --log-file gpurun_out/launches_r02_final.csv $CMD > gpurun_out/ncu_final_list.log 2>&1
echo "launch list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'attn_bwd_tc_kernel' -s 24 -c 4 -o gpurun_out/prof_final_attn_bwd $CMD > gpurun_out/ncu_final_a.log 2>&1
echo "attn_bwd capture rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'attn_fwd_tc_kernel' -s 24 -c 4 -o gpurun_out/prof_final_attn_fwd $CMD > gpurun_out/ncu_final_f.log 2>&1
echo "attn_fwd capture rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'projce_kernel' -s 9 -c 3 -o gpurun_out/prof_final_projce $CMD > gpurun_out/ncu_final_p.log 2>&1
echo "projce capture rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'in_apply_bwd_kernel|dropout_wide_kernel' -s 20 -c 6 -o gpurun_out/prof_final_stream $CMD > gpurun_out/ncu_final_s.log 2>&1
echo "stream capture rc=$?"
ls -la gpurun_out/*final*.ncu-rep gpurun_out/launches_r02_final.csv
