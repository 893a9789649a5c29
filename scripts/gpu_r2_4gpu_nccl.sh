#!/bin/bash
# A/B of the number of NCCL channels at N=4 (the decoder bucket's all-reduce overlaps the encoder backward and takes SMs from it)
set -u
for ch in default 4 8; do
  if [ "$ch" = default ]; then unset NCCL_MAX_NCHANNELS; else export NCCL_MAX_NCHANNELS=$ch; fi
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 12 --warmup 3 --no-cpu --no-library --no-decode --modality-drop 0 > gpurun_out/bench_n4_ch$ch.json 2> gpurun_out/bench_n4_ch$ch.err
  python - $ch <<'PY'
import json, sys
ch = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/bench_n4_ch{ch}.json").read().strip().splitlines()[-1])
    print("channels", ch, "ms", round(d["ms_per_step"], 3), "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1))
except Exception as e:
    print("channels", ch, "ERR", e)
PY
done
