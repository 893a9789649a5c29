#!/bin/bash
set -u
for b in 8 16 32; do
for pf in 0 2400; do
echo "== batch $b pf_cross $pf"
DECODE_BATCH=$b OMR_DECODE_PF_CROSS=$pf OMR_DECODE_PF_MASK=0x70 timeout 200 python scripts/decode_timing.py 300 2>&1 | tail -2
done; done
echo "== batch 16 NB 3"
DECODE_BATCH=16 OMR_DECODE_NB=3 OMR_DECODE_PF_MASK=0x70 timeout 200 python scripts/decode_timing.py 300 2>&1 | tail -2
echo "== batch 8 NB 3"
DECODE_BATCH=8 OMR_DECODE_NB=3 OMR_DECODE_PF_MASK=0x70 timeout 200 python scripts/decode_timing.py 300 2>&1 | tail -2
