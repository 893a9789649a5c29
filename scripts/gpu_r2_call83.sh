#!/bin/bash
# final evidence of the decode work: full GPU suite, ncu full capture of the final kernel (100 steps), default bench line
set -u
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
DCMD="python scripts/decode_timing.py 100 --no-timing"
$DCMD > gpurun_out/plain_dec3.log 2>&1 || { echo plain failed; tail -5 gpurun_out/plain_dec3.log; exit 1; }
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'decode_persistent_kernel' -s 1 -c 1 -o gpurun_out/prof_r2_decode_final2 -f $DCMD > gpurun_out/ncu_dec_final2.log 2>&1
echo "decode capture rc=$?"
t0=$(date +%s)
timeout 900 python bench.py > gpurun_out/bench_r2r.json 2> gpurun_out/bench_r2r.err; echo "bench rc=$? in $(( $(date +%s) - t0 )) s"; tail -2 gpurun_out/bench_r2r.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r2r.json").read().strip().splitlines()[-1])
print("value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "decode", d["decode"] and (round(d["decode"]["value"]), round(d["decode"]["frac_of_hbm_roofline"], 3)), "moddrop", d["modality_drop"]["ms_per_step"])
PY
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
