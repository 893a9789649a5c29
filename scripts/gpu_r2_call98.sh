#!/bin/bash
# the claim "the attention phases do not depend on the batch" on the FINAL kernel: per-phase counters at batch 8 / 16 / 32
set -u
for b in 8 16 32; do echo "== batch $b"; DECODE_BATCH=$b timeout 200 python scripts/decode_timing.py 400 2>&1 | tail -2; done
