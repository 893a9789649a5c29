#!/bin/bash
set -u
for d in 62 126 190 254; do echo "OMR_ATTN_DEBUG=$d"; OMR_ATTN_DEBUG=$d timeout 200 python scripts/bench_attn.py 20 2>&1 | grep "no dropout"; done
python - <<'PY'
import torch, sys
sys.path.insert(0, '.')
# aux costs: memset 16.8 MB, and the small kernels, timed alone
x = torch.empty(32*4*512*65, device='cuda')
def timed(fn, reps=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
print("memset 17MB us", timed(lambda: x.zero_()))
PY
