#!/bin/bash
set -u
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/t_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/t_gpu.log
python scripts/bench_conv.py fwd dgrad wgrad > gpurun_out/conv_uniform.txt 2>&1; tail -30 gpurun_out/conv_uniform.txt
timeout 120 python scripts/bench_attn.py > gpurun_out/bench_attn.txt 2>&1; cat gpurun_out/bench_attn.txt
timeout 300 python bench.py --no-cpu --no-library > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r2b.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_r2b.json").read().strip().splitlines()[-1])
    print("value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "moddrop", d["modality_drop"] and round(d["modality_drop"]["ms_per_step"], 3),
          "decode", d["decode"] and round(d["decode"]["value"]), "roof", d["roofline"]["kernel"], round(d["roofline"]["frac"], 3))
    for k, v in d["breakdown_ms"].items():
        print("  ", k, v)
except Exception as e:
    print("bench parse ERR", e)
PY
