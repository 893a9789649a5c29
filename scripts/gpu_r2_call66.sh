#!/bin/bash
# decode kernel: per-phase cycle counters at the C4 per-GPU size (start of the decode work)
set -u
timeout 300 python scripts/decode_timing.py 2>&1 | tail -5
timeout 300 python scripts/decode_timing.py --no-timing 2>&1 | tail -1
