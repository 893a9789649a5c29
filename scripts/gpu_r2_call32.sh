#!/bin/bash
set -u
for d in 260 268 276 292 316; do echo "OMR_ATTN_DEBUG=$d"; OMR_ATTN_DEBUG=$d timeout 200 python scripts/attn_stamps.py 2>&1 | tail -6 | grep "scores_issued\|grads_issued\|ew_arrived" | cut -c1-190; done
