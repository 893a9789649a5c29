#!/usr/bin/env python
"""Per-entry-point device times of the greedy decode step (eager, CUDA events per C-ABI call) at BASELINE config 4
per-GPU size: batch 32, memory 2337, bf16."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from oracle import synth
import omr_a2s_multimodal_transformer_b200 as pkg
from omr_a2s_multimodal_transformer_b200 import _lib

w2i, i2w = synth.load_vocab()
dev = torch.device("cuda", 0)
m = pkg.MultimodalTransformer(128, 1024, 195, 808, 1268, w2i, i2w).to(dev).eval()
m.set_compute_dtype(torch.bfloat16)
xi, _, xa, _, _, _ = bench.make_batch(32, w2i, seed=500)
with torch.no_grad():
    mem, _ = m._memory(xi.to(dev), xa.to(dev), None, None, "both")
    r = m._decoder_runner()
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 600
    r.decode(mem, w2i["<sos>"], w2i["<eos>"], 0, max_steps=8, stop_at_eos=False, use_graph=False)
    torch.cuda.synchronize()
    _lib.prof_start()
    r.decode(mem, w2i["<sos>"], w2i["<eos>"], 0, max_steps=steps, stop_at_eos=False, use_graph=False)
    prof = _lib.prof_stop()
tot = sum(d["ms"] for d in prof.values())
print(f"decode {steps} steps eager: {tot:.1f} ms in kernels = {tot / steps * 1e3:.1f} us/step")
for k, d in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
    print(f"{k:28s} calls {d['calls']:6d}  {d['ms']:9.2f} ms  {d['ms'] / d['calls'] * 1e3:8.1f} us/call  {d['bytes'] / max(d['ms'], 1e-9) / 1e6:8.1f} GB/s")
