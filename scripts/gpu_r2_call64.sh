#!/bin/bash
set -u
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -m gpu -q -x -k "adam" 2>&1 | tail -3
timeout 300 python bench.py --no-cpu --no-library --no-decode --modality-drop 0 > gpurun_out/bench_adam.json 2>/dev/null
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_adam.json").read().strip().splitlines()[-1])
print("ms", round(d["ms_per_step"], 3))
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'adam_kernel' -c 3 --csv --log-file gpurun_out/adam_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-decode --no-graph --no-library --modality-drop 0 > /dev/null 2>&1
python scripts/summarize_launches.py gpurun_out/adam_launches.csv | tail -2
