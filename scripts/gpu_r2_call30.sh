#!/bin/bash
# full GPU suite + default bench after the attention kernel rewrite (fwd: 8 softmax warps, O in TMEM; bwd: persistent, half-tile pipeline, dQ by TMA reduce)
set -u
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/t_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/t_gpu.log
timeout 400 python bench.py --no-cpu --no-library --no-decode > gpurun_out/bench_r2j.json 2> gpurun_out/bench_r2j.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_r2j.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_r2j.json").read().strip().splitlines()[-1])
    print("value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1))
    print("roof", json.dumps(d["roofline"])[:400])
    for k, v in list(d["breakdown_ms"].items())[:16]: print("  ", k, round(v["ms"], 3), v["calls"])
except Exception as e:
    print("bench parse ERR", e)
PY
