#!/usr/bin/env python
"""Diagnostic: per-tile clock stamps of the halo convolution kernel (OMR_CONV_DEBUG=8 [+1 +2 +4])."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from omr_a2s_multimodal_transformer_b200 import ops
dev = torch.device("cuda", 0)
for (h, w, ci, co) in [(128, 1024, 16, 16), (64, 512, 64, 64)]:
    x = torch.rand(32, h, w, ci, device=dev).to(torch.bfloat16)
    wp = ops.pack_conv_weight(torch.randn(co, ci, 3, 3, device=dev) * 0.05, torch.bfloat16, False)
    b = torch.zeros(co, device=dev)
    ops.conv3x3_fwd(x, wp, b, (1, 1), True)
    torch.cuda.synchronize()

