#!/bin/bash
set -u
for kb in 225 110; do
  OMR_CONV_SMEM_KB=$kb timeout 300 python bench.py --no-cpu --no-library --no-decode --modality-drop 0 > gpurun_out/bench_smem$kb.json 2>/dev/null
  python - <<PY
import json
d = json.loads(open("gpurun_out/bench_smem$kb.json").read().strip().splitlines()[-1])
print("smem cap $kb KB: ms", round(d["ms_per_step"], 3), {k: v["ms"] for k, v in d["breakdown_ms"].items() if "conv3x3" in k})
PY
done
