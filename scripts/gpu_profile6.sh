#!/bin/bash
# end-of-round profile after the stream-overlap changes: full bench line, launch list of eager training steps, and a
# fresh `ncu --set full` capture of the dominant kernel (cross-attention backward, with masked-tile skipping)
set -u
timeout 300 python bench.py > gpurun_out/bench_final3_n1.json 2> gpurun_out/bench_final3_n1.err
echo "bench rc=$?"; tail -c 600 gpurun_out/bench_final3_n1.json
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --no-decode --no-graph --no-library"
timeout 200 $CMD > gpurun_out/plain11.log 2> gpurun_out/plain11.err &&
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 2700 --csv --log-file gpurun_out/launches_r01_final3.csv $CMD > gpurun_out/ncu15.log 2>&1
echo "launch list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'attn_bwd_tc_kernel' -s 16 -c 4 -o gpurun_out/prof_r1_final3 $CMD > gpurun_out/ncu16.log 2>&1
echo "full capture rc=$?"
