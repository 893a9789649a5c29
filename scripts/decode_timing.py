#!/usr/bin/env python
"""Per-phase cycle counters of the persistent decode kernel (OMR_DECODE_TIMING=1) at BASELINE config 4 per-GPU size."""
import os, sys
if "--no-timing" not in sys.argv:
    os.environ["OMR_DECODE_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from oracle import synth
import omr_a2s_multimodal_transformer_b200 as pkg

w2i, i2w = synth.load_vocab()
dev = torch.device("cuda", 0)
m = pkg.MultimodalTransformer(128, 1024, 195, 808, 1268, w2i, i2w).to(dev).eval()
m.set_compute_dtype(torch.bfloat16)
BATCH = int(os.environ.get("DECODE_BATCH", "32"))
xi, _, xa, _, _, _ = bench.make_batch(BATCH, w2i, seed=500)
steps = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 1268
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
print(f"{steps} steps: {e0.elapsed_time(e1):.1f} ms -> {BATCH * steps / e0.elapsed_time(e1) * 1e3:.0f} tokens/s (batch {BATCH})")
