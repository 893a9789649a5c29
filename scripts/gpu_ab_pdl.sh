#!/bin/bash
# GPU tests + A/B of programmatic dependent launch (OMR_PDL) on the training bench
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/t_pdl.log 2>&1; tail -5 gpurun_out/t_pdl.log
for v in 1 0; do
  OMR_PDL=$v timeout 90 python bench.py --no-cpu --no-library --no-decode > gpurun_out/bench33_pdl$v.json 2> gpurun_out/bench33_pdl$v.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench33_pdl$v.json").read().strip().splitlines()[-1])
    print("pdl$v", d["ms_per_step"], d["value"], d["e2e"]["value"])
except Exception as e:
    print("pdl$v ERR", e)
PY
done
tail -3 gpurun_out/bench33_pdl1.err
