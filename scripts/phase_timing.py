#!/usr/bin/env python
"""Where does a C3 training step spend its time?  Eager steps (streams as in the replayed graph) with CUDA events recorded
on the main stream at the phase boundaries: encoders forward (+ PE, concat) | decoder forward + classifier + CE |
classifier / decoder backward (up to dL/dmemory) | encoders backward | gradient sync + Adam.  The phases are bounded by the
points where the main stream joins its side streams, so the figures add up to the eager step time.
Usage: python scripts/phase_timing.py [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from oracle import synth  # noqa: E402
import omr_a2s_multimodal_transformer_b200 as pkg  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
dev = torch.device("cuda", 0)
cfg = bench.CONFIGS["C3"]
w2i, i2w = synth.load_vocab()
torch.manual_seed(0)
model = bench.build_model(cfg, w2i, i2w).to(dev)
model.set_compute_dtype(torch.bfloat16)
model.train()
dp = pkg.DataParallel(model, broadcast=False)
opt = model.configure_optimizers()
opt.grad_scale = dp.grad_scale
batch = [t.to(dev) for t in bench.config_batch(cfg, cfg["batch"], w2i, seed=100)]
stream = torch.cuda.current_stream(dev)


def ev():
    return torch.cuda.Event(enable_timing=True)


def step(rec=None):
    e = [ev() for _ in range(6)] if rec is not None else None
    dp.zero_grad()
    xi, xli, xa, xla, y_in, y_out = batch
    y_in = model.apply_teacher_forcing(y_in)
    if e: e[0].record(stream)
    mem, xl = model._memory(xi, xa, xli, xla, "both")
    if e: e[1].record(stream)
    if e: mem.register_hook(lambda g_: (e[3].record(stream), g_)[1])
    loss = model.decoder.loss(tgt=y_in, memory=mem, memory_len=xl, targets=y_out)
    if e: e[2].record(stream)
    loss.backward()
    if e: e[4].record(stream)
    dp.sync_gradients()
    opt.step()
    if e: e[5].record(stream)
    if rec is not None: rec.append(e)
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
rec = []
for _ in range(reps):
    step(rec)
torch.cuda.synchronize()
names = ["encoders fwd (+PE, concat)", "decoder fwd + classifier + CE", "CE / decoder bwd (to dL/dmem)", "encoders bwd", "grad sync + Adam"]
tot = 0.0
for i, n in enumerate(names):
    ms = sorted(r[i].elapsed_time(r[i + 1]) for r in rec)[len(rec) // 2]
    tot += ms
    print(f"{n:32s} {ms:7.3f} ms")
print(f"{'sum (eager step)':32s} {tot:7.3f} ms")
