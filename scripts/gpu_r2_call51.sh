#!/bin/bash
set -u
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "conv3x3" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_golden.py -m gpu -q -x -k "full_size" 2>&1 | tail -3
python - <<'PY'
import torch, sys
sys.path.insert(0, '.')
from omr_a2s_multimodal_transformer_b200 import ops
dev='cuda:0'
def timed(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
for (n,h,w) in ((32,195,808),(32,128,1024)):
    x = torch.randn(n,h,w,1,device=dev).bfloat16(); dy = torch.randn(n,h,w,16,device=dev).bfloat16()
    dw = torch.zeros(16,1,3,3,device=dev); db = torch.zeros(16,device=dev)
    wp = ops.pack_conv_weight(torch.randn(16,1,3,3,device=dev), torch.bfloat16, False); b = torch.zeros(16, device=dev)
    print(n,h,w,"wgrad us", timed(lambda: ops.conv3x3_wgrad(x, dy, dw, None, (1,1), accumulate=True)), "fwd us", timed(lambda: ops.conv3x3_fwd(x, wp, b, (1,1), relu=True)))
PY
