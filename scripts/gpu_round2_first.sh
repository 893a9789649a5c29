#!/bin/bash
# First GPU call of round 2: validate what was written after round 1's GPU budget was spent.
#   gpurun --timeout 900 -- 'bash scripts/gpu_round2_first.sh'
set -u
timeout 400 python -m pytest tests -m gpu_next -q > gpurun_out/t_gpu_next.log 2>&1; echo "gpu_next rc=$?"; tail -15 gpurun_out/t_gpu_next.log
for v in 0 1; do
  OMR_FUSE_DECODER_LINKS=$v timeout 120 python bench.py --no-cpu --no-library --no-decode > gpurun_out/bench_fuse_$v.json 2> gpurun_out/bench_fuse_$v.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_fuse_$v.json").read().strip().splitlines()[-1])
    print("OMR_FUSE_DECODER_LINKS=$v", round(d["ms_per_step"], 3), "ms/step", round(d["value"], 1), "samples/s, e2e", round(d["e2e"]["value"], 1), "launches/step", d["gpu_launches"] // d["steps"])
except Exception as e:
    print("OMR_FUSE_DECODER_LINKS=$v ERR", e)
PY
done
timeout 120 python bench.py --prefetch --no-cpu --no-library --no-decode > gpurun_out/bench_prefetch.json 2> gpurun_out/bench_prefetch.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_prefetch.json").read().strip().splitlines()[-1])
    print("--prefetch", round(d["value"], 1), "samples/s resident,", round(d["e2e"]["value"], 1), "e2e")
except Exception as e:
    print("--prefetch ERR", e)
PY
timeout 120 python scripts/bench_attn.py > gpurun_out/bench_attn.txt 2>&1; cat gpurun_out/bench_attn.txt
