#!/bin/bash
# decode kernel: ncu full capture with source (NB = 2 default and NB = 3), 100 steps
set -u
DCMD="python scripts/decode_timing.py 100 --no-timing"
$DCMD > gpurun_out/plain_dec.log 2>&1 || { echo plain failed; tail -5 gpurun_out/plain_dec.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:'decode_persistent_kernel' -s 1 -c 1 -o gpurun_out/prof_r2_decode_nb2 -f $DCMD > gpurun_out/ncu_dec.log 2>&1
echo "decode capture rc=$?"
OMR_DECODE_NB=3 ncu --set full --clock-control none --import-source on -k regex:'decode_persistent_kernel' -s 1 -c 1 -o gpurun_out/prof_r2_decode_nb3 -f $DCMD > gpurun_out/ncu_dec3.log 2>&1
echo "decode capture nb3 rc=$?"
OMR_DECODE_DBG_PHASE=1 timeout 300 python scripts/decode_timing.py 400 2>&1 | tail -3
