#!/usr/bin/env python
"""One weight-gradient call at a BASELINE layer shape (for ncu captures): python scripts/wgrad_one.py H W Ci Co sh sw"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from omr_a2s_multimodal_transformer_b200 import ops
dev = torch.device("cuda", 0)
h, w, ci, co, sh, sw = (int(v) for v in (sys.argv[1:7] if len(sys.argv) >= 7 else (16, 128, 128, 128, 1, 1)))
x = torch.randn(32, h, w, ci, device=dev, dtype=torch.bfloat16)
ho, wo = -(-h // sh), -(-w // sw)
dy = torch.randn(32, ho, wo, co, device=dev, dtype=torch.bfloat16)
dw = torch.zeros(co, ci, 3, 3, device=dev)
for _ in range(3):
    ops.conv3x3_wgrad(x, dy, dw, None, (sh, sw), True)
torch.cuda.synchronize()
print("ok", float(dw.abs().mean()))
