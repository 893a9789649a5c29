#!/usr/bin/env python
"""Time the attention entry points on the decoder shapes of BASELINE config 3 (bf16, batch 32, 4 heads x 64):
self-attention (packed qkv, causal, T = 512) and cross-attention (T_q = 512, T_k = 2337 with a ragged -inf key bias),
forward and backward, with and without the attention-probability dropout of train mode.  The iteration loop for the
round-2 work on the dominant kernel.  Usage: python scripts/bench_attn.py [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from omr_a2s_multimodal_transformer_b200 import ops  # noqa: E402
from omr_a2s_multimodal_transformer_b200.ops import AttnSpec  # noqa: E402

dev = torch.device("cuda", 0)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
B, H, HD, T, S = 32, 4, 64, 512, 2337
D = H * HD
g = torch.Generator(device="cpu").manual_seed(0)


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


qkv = (torch.randn(B, T, 3 * D, generator=g) * 0.5).to(dev).to(torch.bfloat16)
q = (torch.randn(B, T, D, generator=g) * 0.5).to(dev).to(torch.bfloat16)
kv = (torch.randn(B, S, 2 * D, generator=g) * 0.5).to(dev).to(torch.bfloat16)
lens = torch.randint(1400, S + 1, (B,), generator=g)
bias = torch.zeros(B, S)
for b, n in enumerate(lens.tolist()):
    bias[b, n:] = float("-inf")
bias = bias.to(dev)

for name, drop in (("no dropout", 0.0), ("dropout 0.1", 0.1)):
    self_spec = AttnSpec(H, HD, causal=True)
    cross_spec = AttnSpec(H, HD, key_bias=bias)
    if drop > 0:
        self_spec = ops.attn_spec_with_dropout(self_spec, drop, 1234)
        cross_spec = ops.attn_spec_with_dropout(cross_spec, drop, 4321)
    # self-attention
    o, lse = ops.attn_fwd(qkv, 0, qkv, D, qkv, 2 * D, self_spec)
    do = torch.randn_like(o)
    dqkv = torch.empty_like(qkv)
    ms_f = timed(lambda: ops.attn_fwd(qkv, 0, qkv, D, qkv, 2 * D, self_spec))
    ms_b = timed(lambda: ops.attn_bwd(qkv, 0, qkv, D, qkv, 2 * D, o, do, lse, dqkv, 0, dqkv, D, dqkv, 2 * D, self_spec))
    fl = 4.0 * B * H * T * T * HD  # full square, as the reference executes it (the kernel skips the masked tiles)
    print(f"self  T={T:5d}          {name:12s} fwd {ms_f * 1e3:7.1f} us {fl / ms_f / 1e9:6.0f} TF/s | "
          f"bwd {ms_b * 1e3:7.1f} us {2.5 * fl / ms_b / 1e9:6.0f} TF/s")
    # cross-attention
    o2, lse2 = ops.attn_fwd(q, 0, kv, 0, kv, D, cross_spec)
    do2 = torch.randn_like(o2)
    dq, dkv = torch.empty_like(q), torch.empty_like(kv)
    ms_f = timed(lambda: ops.attn_fwd(q, 0, kv, 0, kv, D, cross_spec))
    ms_b = timed(lambda: ops.attn_bwd(q, 0, kv, 0, kv, D, o2, do2, lse2, dq, 0, dkv, 0, dkv, D, cross_spec))
    fl = 4.0 * B * H * T * S * HD
    print(f"cross Tq={T:4d} Tk={S:5d} {name:12s} fwd {ms_f * 1e3:7.1f} us {fl / ms_f / 1e9:6.0f} TF/s | "
          f"bwd {ms_b * 1e3:7.1f} us {2.5 * fl / ms_b / 1e9:6.0f} TF/s")
