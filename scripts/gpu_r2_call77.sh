#!/bin/bash
# evidence after the decode rewrite: full ncu capture of the persistent decode kernel (100 steps), phase counters,
# batch-size independence of the attention phases, default bench line
set -u
DCMD="python scripts/decode_timing.py 100 --no-timing"
$DCMD > gpurun_out/plain_dec2.log 2>&1 || { echo plain failed; tail -5 gpurun_out/plain_dec2.log; exit 1; }
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'decode_persistent_kernel' -s 1 -c 1 -o gpurun_out/prof_r2_decode_final -f $DCMD > gpurun_out/ncu_dec_final.log 2>&1
echo "decode capture rc=$?"
for b in 8 32; do DECODE_BATCH=$b timeout 200 python scripts/decode_timing.py 1268 2>&1 | tail -2; done
t0=$(date +%s)
timeout 900 python bench.py > gpurun_out/bench_r2q.json 2> gpurun_out/bench_r2q.err; echo "bench rc=$? in $(( $(date +%s) - t0 )) s"; tail -2 gpurun_out/bench_r2q.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r2q.json").read().strip().splitlines()[-1])
print("value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "decode", d["decode"] and (round(d["decode"]["value"]), round(d["decode"]["frac_of_hbm_roofline"], 3)))
PY
