#!/usr/bin/env python
"""Per-tensor gradient error of the small multimodal fp32 model against the fp64 oracle (debug aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import restate, synth
from tests.helpers import build_multimodal, oracle_truth_and_floors
DEV = "cuda:0"
m, sd, w2i = build_multimodal(mixer="concat", dtype=torch.float32)
xi, xli, xa, xla, y_in, y_out = synth.synth_multimodal_batch(3, (64, 128), (48, 96), [20, 12, 7], w2i)
truth = oracle_truth_and_floors(lambda s, dt: restate.multimodal_forward(s, xi, xli, xa, xla, y_in, mixer_type="concat", dtype=dt), y_out, sd)
print("floors", truth["floor_logits_fp32"], truth["floor_grads_fp32"])
for rep in range(2):
    m.zero_grad(set_to_none=True)
    mem, xl = m._memory(xi.to(DEV), xa.to(DEV), xli.to(DEV), xla.to(DEV), "both")
    loss = m.decoder.loss(tgt=y_in.to(DEV), memory=mem, memory_len=xl, targets=y_out.to(DEV))
    loss.backward()
    rows, num, den = [], 0.0, 0.0
    for k, p in m.named_parameters():
        if k not in truth["grads"] or p.grad is None:
            continue
        a, b = p.grad.detach().double().cpu().reshape(-1), truth["grads"][k].double().reshape(-1)
        e2 = float((a - b).pow(2).sum())
        num += e2; den += float(b.pow(2).sum())
        rows.append((e2, k, float(b.norm())))
    rows.sort(reverse=True)
    print("global_rel", (num / den) ** 0.5)
    for e2, k, bn in rows[:6]:
        print(f"   {k:55s} abs_err {e2 ** 0.5:.3e}  |ref| {bn:.3e}  share {e2 / num:.2f}")
    if rep == 0:
        for k, p in m.named_parameters():
            if k.startswith("audio_encoder") and k in truth["grads"]:
                a, b = p.grad.detach().double().cpu().reshape(-1), truth["grads"][k].double().reshape(-1)
                print(f"      {k:50s} rel {float((a - b).norm() / (b.norm() + 1e-30)):.2e}")
