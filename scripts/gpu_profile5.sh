#!/bin/bash
# round-1 final profile (after the last kernel changes): launch list of eager training steps + full ncu sections of the
# kernels that lead the step + the persistent decode kernel
set -u
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --no-decode --no-graph --no-library"
$CMD > gpurun_out/plain9.log 2> gpurun_out/plain9.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 2700 --csv --log-file gpurun_out/launches_r01_final2.csv $CMD > gpurun_out/ncu12.log 2>&1
echo "launch list rc=$?"
# one training step ~ 885 launches; skip the first (warm-up) step and take the kernels of interest from the second
ncu --set full --clock-control none --import-source on -k regex:'attn_bwd_tc_kernel|attn_fwd_tc_kernel|wgrad_small_kernel|wgrad_tc_kernel|gemm_tc_kernel' -s 330 -c 28 -o gpurun_out/prof_r1_final2 $CMD > gpurun_out/ncu13.log 2>&1
echo "full capture rc=$?"
DCMD="python scripts/decode_timing.py 48 --no-timing"
$DCMD > gpurun_out/plain10.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'decode_persistent_kernel' -c 2 -o gpurun_out/prof_r1_decode2 $DCMD > gpurun_out/ncu14.log 2>&1
echo "decode capture rc=$?"
