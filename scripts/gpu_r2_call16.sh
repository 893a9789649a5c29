#!/bin/bash
set -u
echo "== wgrad_small on (default)"; python scripts/bench_conv.py wgrad 2>&1 | tail -25
echo "== OMR_WGRAD_SMALL=0 (tcgen05 wgrad everywhere)"; OMR_WGRAD_SMALL=0 python scripts/bench_conv.py wgrad 2>&1 | tail -25
