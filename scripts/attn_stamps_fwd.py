"""debug aid: run the cross-attention forward a few times so that OMR_ATTN_DEBUG=512 prints CTA (0,0,0)'s clock stamps"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from omr_a2s_multimodal_transformer_b200 import ops
from omr_a2s_multimodal_transformer_b200.ops import AttnSpec
dev = torch.device("cuda", 0)
B, H, HD, T, S = 32, 4, 64, 512, 2337
D = H * HD
g = torch.Generator(device="cpu").manual_seed(0)
q = (torch.randn(B, T, D, generator=g) * 0.5).to(dev).to(torch.bfloat16)
kv = (torch.randn(B, S, 2 * D, generator=g) * 0.5).to(dev).to(torch.bfloat16)
bias = torch.zeros(B, S)
lens = torch.randint(1400, S + 1, (B,), generator=g)
for b, n in enumerate(lens.tolist()):
    bias[b, n:] = float("-inf")
bias = bias.to(dev)
spec = AttnSpec(H, HD, key_bias=bias)
if len(sys.argv) > 1 and sys.argv[1] == "drop":
    spec = ops.attn_spec_with_dropout(spec, 0.1, 4321)
for _ in range(6):
    ops.attn_fwd(q, 0, kv, 0, kv, D, spec)
torch.cuda.synchronize()
