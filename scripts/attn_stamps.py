"""debug aid: run the cross-attention backward a few times so that OMR_ATTN_DEBUG=256 prints CTA 0's clock stamps"""
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
lens = torch.randint(1400, S + 1, (B,), generator=g)
bias = torch.zeros(B, S)
for b, n in enumerate(lens.tolist()):
    bias[b, n:] = float("-inf")
bias = bias.to(dev)
spec = AttnSpec(H, HD, key_bias=bias)
if len(sys.argv) > 1 and sys.argv[1] == "drop":
    spec = ops.attn_spec_with_dropout(spec, 0.1, 4321)
o2, lse2 = ops.attn_fwd(q, 0, kv, 0, kv, D, spec)
do2 = torch.randn_like(o2)
dq, dkv = torch.empty_like(q), torch.empty_like(kv)
for _ in range(6):
    ops.attn_bwd(q, 0, kv, 0, kv, D, o2, do2, lse2, dq, 0, dkv, 0, dkv, D, spec)
torch.cuda.synchronize()
