#!/bin/bash
# full GPU suite + default bench line after the decode rewrite
set -u
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
t0=$(date +%s)
timeout 900 python bench.py > gpurun_out/bench_r2p.json 2> gpurun_out/bench_r2p.err; echo "bench rc=$? in $(( $(date +%s) - t0 )) s"; tail -2 gpurun_out/bench_r2p.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r2p.json").read().strip().splitlines()[-1])
print("value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "decode", d["decode"] and (round(d["decode"]["value"]), round(d["decode"]["frac_of_hbm_roofline"], 3)))
PY
