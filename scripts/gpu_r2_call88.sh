#!/bin/bash
# the default bench line and the reference arm with the final bench.py
set -u
t0=$(date +%s)
timeout 900 python bench.py > gpurun_out/bench_r2s.json 2> gpurun_out/bench_r2s.err; echo "bench rc=$? in $(( $(date +%s) - t0 )) s"; tail -2 gpurun_out/bench_r2s.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r2s.json").read().strip().splitlines()[-1])
print("value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "decode", d["decode"] and (round(d["decode"]["value"]), round(d["decode"]["frac_of_hbm_roofline"], 3)), "moddrop", d["modality_drop"]["ms_per_step"], "cpu", d["cpu_baseline"]["value"])
print({k: d["roofline"][k] for k in ("kernel", "achieved", "frac", "traffic", "avg_ms", "share_of_step")})
PY
t0=$(date +%s)
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2s_ref.json 2> gpurun_out/bench_r2s_ref.err; echo "ref arm rc=$? in $(( $(date +%s) - t0 )) s"; tail -c 300 gpurun_out/bench_r2s_ref.json
