#!/usr/bin/env python
"""One forward convolution at a BASELINE layer shape (for ncu captures): python scripts/conv_one.py [Ci Co [H W]]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from omr_a2s_multimodal_transformer_b200 import ops
dev = torch.device("cuda", 0)
ci = int(sys.argv[1]) if len(sys.argv) > 1 else 16
co = int(sys.argv[2]) if len(sys.argv) > 2 else 16
h = int(sys.argv[3]) if len(sys.argv) > 3 else 128
w = int(sys.argv[4]) if len(sys.argv) > 4 else 1024
x = torch.rand(32, h, w, ci, device=dev).to(torch.bfloat16)
wp = ops.pack_conv_weight(torch.randn(co, ci, 3, 3, device=dev) * 0.05, torch.bfloat16, False)
b = torch.zeros(co, device=dev)
for _ in range(3):
    y = ops.conv3x3_fwd(x, wp, b, (1, 1), True)
torch.cuda.synchronize()
print("ok", float(y.float().abs().mean()))
