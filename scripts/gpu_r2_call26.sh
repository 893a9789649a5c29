#!/bin/bash
set -u
for d in 256 510; do echo "OMR_ATTN_DEBUG=$d"; OMR_ATTN_DEBUG=$d timeout 200 python scripts/attn_stamps.py 2>&1 | tail -8; done
