#!/bin/bash
set -u
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/t_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/t_gpu.log
timeout 400 python bench.py > gpurun_out/bench_r2i.json 2> gpurun_out/bench_r2i.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_r2i.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_r2i.json").read().strip().splitlines()[-1])
    print("value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "moddrop", d["modality_drop"] and round(d["modality_drop"]["ms_per_step"], 3),
          "decode", d["decode"] and round(d["decode"]["value"]), "lib", d["library_baseline"], "cpu", d["cpu_baseline"])
    print("roof", json.dumps(d["roofline"])[:600])
    for k in d["top_kernels"]: print("  ", k)
except Exception as e:
    print("bench parse ERR", e)
PY
for c in C2 C5 C5long; do
  timeout 400 python bench.py --config $c --no-cpu --no-library > gpurun_out/bench_r2_$c.json 2> gpurun_out/bench_r2_$c.err; echo "bench $c rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_r2_$c.json").read().strip().splitlines()[-1])
    print("$c value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "tc_frac", round(d["model_tc_frac_of_sustained_peak"], 3), "decode", d["decode"] and round(d["decode"]["value"]))
except Exception as e:
    print("$c parse ERR", e)
PY
done
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2_ref.json 2> gpurun_out/bench_r2_ref.err; echo "ref arm rc=$?"; tail -c 600 gpurun_out/bench_r2_ref.json
