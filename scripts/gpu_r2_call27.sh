#!/bin/bash
set -u
for d in 256 510; do echo "OMR_ATTN_DEBUG=$d"; OMR_ATTN_DEBUG=$d timeout 200 python scripts/attn_stamps.py 2>&1 | tail -8 | cut -c1-200; done
for d in 0 2 62; do echo "OMR_ATTN_DEBUG=$d"; OMR_ATTN_DEBUG=$d timeout 200 python scripts/bench_attn.py 20 2>&1 | grep "self\|cross"; done
