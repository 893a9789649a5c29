#!/bin/bash
set -u
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv
python scripts/decode_timing.py --no-timing 2>&1 | tail -2
python scripts/decode_timing.py --no-timing 2>&1 | tail -2
timeout 300 python bench.py --no-cpu --no-library --modality-drop 0 --variants 2 > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r2c.json").read().strip().splitlines()[-1])
print("variants=2 nodrop: value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "decode", d["decode"] and round(d["decode"]["value"]), d["decode"]["ms"])
PY
timeout 300 python bench.py --no-cpu --no-library > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r2d.json").read().strip().splitlines()[-1])
print("default: value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "decode", d["decode"] and round(d["decode"]["value"]), d["decode"]["ms"], d["clocks"])
PY
