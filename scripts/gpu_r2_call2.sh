#!/bin/bash
# round 2, call 2: conv microbench with the halo kernel's diagnostic switches, full -m gpu suite, rewritten bench.py
set -u
python scripts/bench_conv.py > gpurun_out/conv_base.txt 2>&1; tail -40 gpurun_out/conv_base.txt
for d in 1 2 4 3 7; do
  echo "== OMR_CONV_DEBUG=$d (fwd only, few shapes)"
  BENCH_CONV_FEW=1 OMR_CONV_DEBUG=$d timeout 120 python scripts/bench_conv.py fwd 2>&1 | tail -8
done
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/t_gpu.log
timeout 300 python bench.py --no-cpu --no-library > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r2a.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_r2a.json").read().strip().splitlines()[-1])
    print("value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "moddrop", d["modality_drop"] and round(d["modality_drop"]["ms_per_step"], 3),
          "decode", d["decode"] and round(d["decode"]["value"]), "roof", d["roofline"]["kernel"], round(d["roofline"]["frac"], 3))
    for k, v in d["breakdown_ms"].items():
        print("  ", k, v)
except Exception as e:
    print("bench parse ERR", e)
PY
