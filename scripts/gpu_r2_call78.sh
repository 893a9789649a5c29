#!/bin/bash
# run-to-run spread of the decode leg: the script alone (x3) and inside bench.py (x3, short training legs)
set -u
for i in 1 2 3; do timeout 200 python scripts/decode_timing.py 1268 --no-timing 2>&1 | tail -1; done
for i in 1 2 3; do
timeout 600 python bench.py --no-cpu --no-library --steps 4 --warmup 3 --modality-drop 0 > gpurun_out/bench_spread_$i.json 2> gpurun_out/bench_spread_$i.err
python - $i <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/bench_spread_{sys.argv[1]}.json").read().strip().splitlines()[-1])
print("bench", sys.argv[1], "ms/step", round(d["ms_per_step"], 3), "decode ms", round(d["decode"]["ms"], 1), round(d["decode"]["value"]))
PY
done
