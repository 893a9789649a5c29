#!/bin/bash
set -u
OMR_BENCH_TABLE=gpurun_out/kernel_table.txt timeout 400 python bench.py --no-cpu --no-library --no-decode --modality-drop 0 --steps 4 > gpurun_out/bench_tab.json 2> gpurun_out/bench_tab.err; echo "bench rc=$?"
head -70 gpurun_out/kernel_table.txt
