#!/bin/bash
set -u
timeout 300 python scripts/phase_timing.py 6 2>&1 | tail -8
