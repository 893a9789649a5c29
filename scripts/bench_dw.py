#!/usr/bin/env python
"""Time the depthwise 3x3 entry points at the encoder's DSC shapes (bf16).  Usage: python scripts/bench_dw.py [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from omr_a2s_multimodal_transformer_b200 import ops
dev = torch.device("cuda", 0)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
def timed(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
for (n, h, w, c) in ((32, 8, 128, 128), (32, 8, 128, 256), (32, 13, 101, 128), (32, 13, 101, 256)):
    x = torch.randn(n, h, w, c, device=dev).bfloat16()
    dy = torch.randn(n, h, w, c, device=dev).bfloat16()
    wp = torch.randn(3, 3, c, device=dev).bfloat16()
    bias = torch.randn(c, device=dev)
    dw = torch.zeros(c, 1, 3, 3, device=dev); db = torch.zeros(c, device=dev)
    mb = x.numel() * 2 / 1e6
    tf = timed(lambda: ops.dwconv3x3_fwd(x, wp, bias))
    tw = timed(lambda: ops.dwconv3x3_wgrad(x, dy, dw, db, accumulate=True))
    print(f"{n}x{h}x{w}x{c}: {mb:5.1f} MB/tensor | fwd {tf:6.1f} us ({2*mb/tf:5.2f} TB/s) | wgrad {tw:6.1f} us ({2*mb/tw:5.2f} TB/s)")
