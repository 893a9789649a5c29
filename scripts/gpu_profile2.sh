#!/bin/bash
set -u
CMD="python bench.py --steps 1 --warmup 3 --no-cpu --no-decode"
python scripts/profile_decode.py 300 > gpurun_out/decode_prof.txt 2>&1; tail -20 gpurun_out/decode_prof.txt
$CMD > gpurun_out/plain3.log 2> gpurun_out/plain3.err &&
ncu --set full --clock-control none --import-source on -k regex:'wgrad_tc_kernel|attn_bwd_tc_kernel|attn_fwd_tc_kernel|conv_halo_kernel|conv_tc_kernel' -s 60 -c 14 -o gpurun_out/prof_r1b $CMD > gpurun_out/ncu3.log 2>&1
echo "full capture rc=$?"
