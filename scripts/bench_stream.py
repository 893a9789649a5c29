#!/usr/bin/env python
"""Time the HBM-bound element-wise passes of the encoder at their largest shapes (bf16): InstanceNorm apply forward /
backward (statistics supplied, as in the model) and MixDropout.  Usage: python scripts/bench_stream.py [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from omr_a2s_multimodal_transformer_b200 import ops
dev = torch.device("cuda", 0)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
def timed(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
for (n, h, w, c) in ((32, 195, 808, 32), (32, 128, 1024, 16), (32, 98, 404, 64)):
    x = torch.randn(n, h, w, c, device=dev).bfloat16()
    dy = torch.randn(n, h, w, c, device=dev).bfloat16()
    mb = x.numel() * 2 / 1e6
    sums = torch.zeros(n, c, 2, dtype=torch.float64, device=dev)
    sums[:, :, 0] = x.double().sum((1, 2)); sums[:, :, 1] = (x.double() ** 2).sum((1, 2))
    y, stats = ops.instnorm_fwd(x, 1e-5, sums=sums)
    bs = torch.zeros(n, c, 2, dtype=torch.float64, device=dev)
    bs[:, :, 0] = dy.double().sum((1, 2)); bs[:, :, 1] = (dy.double() * x.double()).sum((1, 2))
    tf = timed(lambda: ops.instnorm_fwd(x, 1e-5, sums=sums))
    tb = timed(lambda: ops.instnorm_bwd(dy, x, stats, relu_mask=True, mask_scale=2.0, sums=bs))
    td = timed(lambda: ops.dropout(x, 0.5, 1234))
    tdc = timed(lambda: ops.dropout(x, 0.25, 1234, channelwise=True))
    print(f"{n}x{h}x{w}x{c} ({mb:.0f} MB): IN fwd {tf:6.1f} us {2*mb/tf:4.2f} TB/s | IN bwd {tb:6.1f} us {3*mb/tb:4.2f} TB/s | "
          f"dropout {td:6.1f} us {2*mb/td:4.2f} TB/s | dropout2d {tdc:6.1f} us {2*mb/tdc:4.2f} TB/s")
