#!/bin/bash
set -u
OMR_ATTN_DEBUG=512 timeout 100 python scripts/attn_stamps_fwd.py 2>&1 | tail -6 | cut -c1-170
OMR_ATTN_DEBUG=512 timeout 100 python scripts/attn_stamps_fwd.py drop 2>&1 | tail -6 | cut -c1-170
