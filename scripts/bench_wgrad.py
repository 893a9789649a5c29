#!/usr/bin/env python
"""Time the 3x3 weight-gradient entry point on the encoder shapes of BASELINE config 3 (bf16, batch 32)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from omr_a2s_multimodal_transformer_b200 import ops

dev = torch.device("cuda", 0)
shapes = [(195, 808, 16, 16, 1, 1), (128, 1024, 16, 16, 1, 1), (195, 808, 16, 32, 1, 1), (195, 808, 32, 32, 1, 1),
          (195, 808, 32, 32, 2, 2), (98, 404, 32, 64, 1, 1), (98, 404, 64, 64, 1, 1), (98, 404, 64, 64, 2, 2),
          (49, 202, 64, 128, 1, 1), (49, 202, 128, 128, 1, 1), (25, 101, 128, 128, 1, 1)]
tot = 0.0
for (h, w, ci, co, sh, sw) in shapes:
    x = torch.randn(32, h, w, ci, device=dev, dtype=torch.bfloat16)
    ho, wo = -(-h // sh), -(-w // sw)
    dy = torch.randn(32, ho, wo, co, device=dev, dtype=torch.bfloat16)
    dw = torch.zeros(co, ci, 3, 3, device=dev)
    db = torch.zeros(co, device=dev)
    for _ in range(2):
        ops.conv3x3_wgrad(x, dy, dw, db, (sh, sw), True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.conv3x3_wgrad(x, dy, dw, db, (sh, sw), True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    tot += ms
    fl = 2.0 * 32 * ho * wo * 9 * ci * co
    by = 2.0 * 32 * (h * w * ci + ho * wo * co)
    print(f"{h}x{w} {ci}->{co} s{sh}: {ms*1e3:7.1f} us  {fl/ms/1e9:7.1f} TF/s  {by/ms/1e6:7.0f} GB/s")
print(f"total {tot:.3f} ms")
