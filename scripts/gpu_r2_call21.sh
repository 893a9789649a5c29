#!/bin/bash
set -u
for d in 0 6 14 22 30 38 62; do echo "OMR_ATTN_DEBUG=$d"; OMR_ATTN_DEBUG=$d timeout 200 python scripts/bench_attn.py 20 2>&1 | grep "no dropout"; done
