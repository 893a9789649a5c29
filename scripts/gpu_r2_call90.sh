#!/bin/bash
# last evidence run of the round: full GPU suite, default bench line, smoke
set -u
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
t0=$(date +%s)
timeout 900 python bench.py > gpurun_out/bench_r2t.json 2> gpurun_out/bench_r2t.err; echo "bench rc=$? in $(( $(date +%s) - t0 )) s"; tail -2 gpurun_out/bench_r2t.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r2t.json").read().strip().splitlines()[-1])
print("value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "decode", d["decode"] and (round(d["decode"]["value"]), round(d["decode"]["frac_of_hbm_roofline"], 3)), "moddrop", d["modality_drop"]["ms_per_step"], "cpu", d["cpu_baseline"]["value"])
PY
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
SWEEP_CFGS='[{}, {"OMR_DECODE_PF_MASK": "0x7f"}, {"OMR_DECODE_PF_MASK": "0x78"}, {"OMR_DECODE_PF_MASK": "0x7c"}, {"OMR_DECODE_PF_SELF": 1024}, {}]' timeout 400 python scripts/decode_sweep2.py 1268 2>&1 | grep -v Warn | tail -6
