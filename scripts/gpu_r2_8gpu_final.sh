#!/bin/bash
# 8-GPU bench on the final code
set -u
t0=$(date +%s)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 8 --warmup 3 --no-cpu --no-library > gpurun_out/bench_final7_n8.json 2> gpurun_out/bench_final7_n8.err; echo "bench n4 rc=$? in $(( $(date +%s) - t0 )) s"
tail -2 gpurun_out/bench_final7_n8.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_final7_n8.json").read().strip().splitlines()[-1])
    print("N=8 value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "moddrop", d["modality_drop"] and round(d["modality_drop"]["ms_per_step"], 3), "decode", d["decode"] and round(d["decode"]["value"]))
except Exception as e:
    print("bench parse ERR", e)
PY
