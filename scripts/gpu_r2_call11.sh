#!/bin/bash
set -u
for st in 0 2; do for d in 8 15; do echo "== STAGE=$st DEBUG=$d"; OMR_CONV_STAGE=$st OMR_CONV_DEBUG=$d python scripts/conv_stamps.py 2>&1 | tail -28; done; done
