#!/bin/bash
# round-1 final profile: launch list of one eager training step, full ncu sections of the top kernels, the decode kernel
set -u
CMD="python bench.py --steps 1 --warmup 3 --no-cpu --no-decode --no-graph --no-library"
$CMD > gpurun_out/plain7.log 2> gpurun_out/plain7.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/launches_r01_final.csv $CMD > gpurun_out/ncu9.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'wgrad_tc_kernel|attn_bwd_tc_kernel|attn_fwd_tc_kernel|conv_halo_kernel|in_apply_bwd_kernel' -s 120 -c 16 -o gpurun_out/prof_r1_final $CMD > gpurun_out/ncu10.log 2>&1
echo "full capture rc=$?"
DCMD="python scripts/decode_timing.py 48 --no-timing"
$DCMD > gpurun_out/plain8.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'decode_persistent_kernel' -c 2 -o gpurun_out/prof_r1_decode $DCMD > gpurun_out/ncu11.log 2>&1
echo "decode capture rc=$?"
ls -la gpurun_out/*.ncu-rep
