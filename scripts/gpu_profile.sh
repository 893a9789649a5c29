#!/bin/bash
# Run on the GPU box (under gpurun): plain bench first, then the ncu launch list and one full capture of the top kernels.
set -u
CMD="python bench.py --steps 1 --warmup 3 --no-cpu --no-decode"
$CMD > gpurun_out/plain.log 2> gpurun_out/plain.err || { echo "plain run failed"; tail -5 gpurun_out/plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1100 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2.log 2> gpurun_out/plain2.err &&
ncu --set full --clock-control none --import-source on -k regex:'wgrad_tc_kernel|attn_bwd_tc_kernel|conv_halo_kernel|gemm_tc_kernel' -s 40 -c 8 -o gpurun_out/prof_r1 $CMD > gpurun_out/ncu2.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out | tail -8
