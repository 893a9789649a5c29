#!/bin/bash
set -u
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_ops.py -q -x > gpurun_out/t_tc.log 2>&1; echo "pytest tc/ops rc=$?"; tail -5 gpurun_out/t_tc.log
python scripts/bench_conv.py fwd dgrad > gpurun_out/conv_v3.txt 2>&1; tail -25 gpurun_out/conv_v3.txt
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_golden.py::test_greedy_long_c1_all_four_samples_to_max_len > gpurun_out/t_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/t_gpu.log
timeout 300 python bench.py --no-cpu --no-library --no-decode > gpurun_out/bench_r2h.json 2> gpurun_out/bench_r2h.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r2h.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_r2h.json").read().strip().splitlines()[-1])
    print("value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "moddrop", d["modality_drop"] and round(d["modality_drop"]["ms_per_step"], 3))
    for k, v in d["breakdown_ms"].items():
        print("  ", k, v)
except Exception as e:
    print("bench parse ERR", e)
PY
