#!/bin/bash
# A/B of programmatic dependent launch on the final step (it was neutral at 25.7 ms/step in round 1)
set -u
for v in 0 1 0 1; do
  OMR_PDL=$v timeout 200 python bench.py --no-cpu --no-library --no-decode --modality-drop 0 > gpurun_out/bench_pdl$v.json 2> gpurun_out/bench_pdl$v.err
  python - $v <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/bench_pdl{v}.json").read().strip().splitlines()[-1])
    print("pdl", v, "ms", round(d["ms_per_step"], 3), "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1))
except Exception as e:
    print("pdl", v, "ERR", e)
PY
done
