#!/usr/bin/env python
"""Sweep of the persistent decode kernel's tuning switches (prefetch points / amounts, start stagger) at the C4
per-GPU size.  Usage: python scripts/decode_sweep2.py [steps] -- one line per setting."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from oracle import synth
import omr_a2s_multimodal_transformer_b200 as pkg

w2i, i2w = synth.load_vocab()
dev = torch.device("cuda", 0)
m = pkg.MultimodalTransformer(128, 1024, 195, 808, 1268, w2i, i2w).to(dev).eval()
m.set_compute_dtype(torch.bfloat16)
xi, _, xa, _, _, _ = bench.make_batch(32, w2i, seed=500)
steps = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 400
KEYS = ["OMR_DECODE_PF_MASK", "OMR_DECODE_PF_CROSS", "OMR_DECODE_PF_SELF", "OMR_DECODE_STAGGER_NS"]


def run(cfg, timing=False):
    for k in KEYS:
        os.environ.pop(k, None)
    os.environ.update({k: str(v) for k, v in cfg.items()})
    with torch.no_grad():
        mem, _ = m._memory(xi.to(dev), xa.to(dev), None, None, "both")
        r = m._decoder_runner()
        r.decode(mem, w2i["<sos>"], w2i["<eos>"], 0, max_steps=8, stop_at_eos=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r.decode(mem, w2i["<sos>"], w2i["<eos>"], 0, max_steps=steps, stop_at_eos=False)
        e1.record()
        torch.cuda.synchronize()
        toks = r.last_tokens.clone() if hasattr(r, "last_tokens") else None
    ms = e0.elapsed_time(e1)
    print(f"{cfg}: {steps} steps {ms:.1f} ms -> {32 * steps / ms * 1e3:.0f} tok/s ({ms / steps * 1e3:.1f} us/step)", flush=True)


cfgs = [
    {},
    {"OMR_DECODE_PF_CROSS": 0},
    {"OMR_DECODE_PF_MASK": "0x7f"},
    {"OMR_DECODE_PF_MASK": "0x78"},
    {"OMR_DECODE_PF_MASK": "0x20"},
    {"OMR_DECODE_PF_SELF": 1024},
    {"OMR_DECODE_STAGGER_NS": 2000},
    {},
]
if os.environ.get("SWEEP_CFGS"):
    import json
    cfgs = json.loads(os.environ["SWEEP_CFGS"])
for c in cfgs:
    run(c)
