#!/bin/bash
# sweep the L2 prefetch depth of the persistent decode kernel (rows of the next attention's K/V stream)
for c in 0 512 1024 2400; do for s in 0 4096; do
echo "pf_cross=$c pf_self=$s: $(OMR_DECODE_PF_CROSS=$c OMR_DECODE_PF_SELF=$s timeout 200 python scripts/decode_timing.py --no-timing 2>&1 | tail -1)"
done; done
OMR_DECODE_PF_CROSS=1024 OMR_DECODE_PF_SELF=4096 timeout 200 python scripts/decode_timing.py 2>&1 | tail -4
