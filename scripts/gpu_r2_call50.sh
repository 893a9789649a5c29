#!/bin/bash
# final record: the default bench line (all legs) and the reference arm
set -u
t0=$(date +%s)
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$? in $(( $(date +%s) - t0 )) s"; tail -2 gpurun_out/bench_final.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_final.json").read().strip().splitlines()[-1])
print("value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "moddrop", d["modality_drop"] and round(d["modality_drop"]["ms_per_step"], 3),
      "decode", d["decode"] and round(d["decode"]["value"]), "lib", d["library_baseline"] and round(d["library_baseline"]["value"], 1), "cpu", d["cpu_baseline"] and d["cpu_baseline"]["value"], "launches", d["gpu_launches"])
print("roof", {k: d["roofline"][k] for k in ("kernel", "achieved", "frac", "traffic", "avg_ms", "share_of_step")})
print("clocks", d["clocks"], "tc_frac", d["model_tc_frac_of_sustained_peak"])
PY
t0=$(date +%s)
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err; echo "ref arm rc=$? in $(( $(date +%s) - t0 )) s"; tail -c 500 gpurun_out/bench_final_ref.json
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -4
