#!/bin/bash
# closing check of the committed state: decode tests, default bench line
set -u
timeout 600 python -m pytest tests -m gpu -x -q -k "greedy or decode or weighted or late or abi" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_r2u.json 2> gpurun_out/bench_r2u.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r2u.json").read().strip().splitlines()[-1])
print("value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "decode", d["decode"] and (round(d["decode"]["value"]), round(d["decode"]["frac_of_hbm_roofline"], 3)))
PY
