#!/bin/bash
set -u
for c in C2 C5 C5long; do
  timeout 400 python bench.py --config $c --no-cpu --no-library > gpurun_out/bench_final_$c.json 2> gpurun_out/bench_final_$c.err; echo "bench $c rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_final_$c.json").read().strip().splitlines()[-1])
    print("$c value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "tc_frac", round(d["model_tc_frac_of_sustained_peak"], 3), "decode", d["decode"] and round(d["decode"]["value"]))
except Exception as e:
    print("$c parse ERR", e)
PY
done
