#!/bin/bash
# 2-GPU checks on the final code: NCCL equivalence test + the bench at N=2
set -u
timeout 600 python -m pytest tests/test_gpu_ddp_nccl.py -q -x > gpurun_out/t_ddp2.log 2>&1; echo "ddp test rc=$?"; tail -3 gpurun_out/t_ddp2.log
t0=$(date +%s)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 --no-cpu --no-library > gpurun_out/bench_final3_n2.json 2> gpurun_out/bench_final3_n2.err; echo "bench n2 rc=$? in $(( $(date +%s) - t0 )) s"
tail -3 gpurun_out/bench_final3_n2.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_final3_n2.json").read().strip().splitlines()[-1])
    print("N=2 value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "moddrop", d["modality_drop"] and round(d["modality_drop"]["ms_per_step"], 3), "decode", d["decode"] and round(d["decode"]["value"]))
except Exception as e:
    print("bench parse ERR", e)
PY
