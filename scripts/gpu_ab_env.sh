#!/bin/bash
# usage: gpu_ab_env.sh VAR tag   -- GPU tests, then the training bench with VAR=1 and VAR=0
VAR=$1; TAG=$2
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/t_$TAG.log 2>&1; tail -4 gpurun_out/t_$TAG.log
for v in 1 0; do
  env $VAR=$v timeout 120 python bench.py --no-cpu --no-library --no-decode > gpurun_out/bench_${TAG}_$v.json 2> gpurun_out/bench_${TAG}_$v.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_${TAG}_$v.json").read().strip().splitlines()[-1])
    print("$VAR=$v", round(d["ms_per_step"], 3), round(d["value"], 1), round(d["e2e"]["value"], 1))
except Exception as e:
    print("$VAR=$v ERR", e)
PY
done
tail -3 gpurun_out/bench_${TAG}_1.err
