#!/bin/bash
set -u
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/t_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t_gpu.log
timeout 400 python bench.py --no-cpu --no-library --no-decode > gpurun_out/bench_r2n.json 2> gpurun_out/bench_r2n.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r2n.json").read().strip().splitlines()[-1])
print("value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "moddrop", d["modality_drop"] and round(d["modality_drop"]["ms_per_step"], 3))
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'ln_bwd_drop|drop_add_ln|colsum' -c 60 --csv --log-file gpurun_out/ln_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-decode --no-graph --no-library --modality-drop 0 > /dev/null 2>&1
python scripts/summarize_launches.py gpurun_out/ln_launches.csv
