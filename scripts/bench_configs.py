#!/usr/bin/env python
"""Training throughput of the other BASELINE.json training configurations on one GPU (bench.py times config 3):
  C2  audio-only A2S model, 1x195x808 spectrograms, batch 16, T = 512, bf16
  C5  scaled-up multimodal model: d_model 512, 8 heads, ff 512, 8 layers, max_len 2536, encoders ending in 512
      channels, image 1x128x1024 + audio 1x195x808 (S = 2337), batch 32, T = 1024 (c5) or 2535 (c5long), bf16
Same protocol as bench.py: full step (zero grads, forward, fused projection + cross-entropy, backward, fused Adam) in
train mode, replayed as a CUDA graph, CUDA-event timing after warm-up.  One JSON line per configuration."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (make_batch, peaks)
from oracle import synth  # noqa: E402


def run(name, model, batch_fn, step_builder, flops_per_sample, b, steps=6, warmup=3):
    import omr_a2s_multimodal_transformer_b200 as pkg

    dev = torch.device("cuda:0")
    model = model.to(dev)
    model.set_compute_dtype(torch.bfloat16)
    model.train()
    dp = pkg.DataParallel(model, broadcast=False)
    opt = model.configure_optimizers()
    opt.grad_scale = dp.grad_scale
    resident = [t.to(dev) for t in batch_fn()]
    step = step_builder(model, dp, opt)
    for _ in range(2):
        step(resident)
    torch.cuda.synchronize()
    stepper = pkg.GraphedTrainStep(step, resident, opt, variants=2, warmup=1)
    for _ in range(warmup):
        stepper()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        loss = stepper()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    pk = bench.peaks()
    tf = flops_per_sample * b / (ms * 1e-3) / 1e12
    out = {"config": name, "metric": "train_samples_per_s", "value": b / (ms * 1e-3), "unit": "samples/s", "ms_per_step": ms,
           "batch": b, "dtype": "bf16", "steps": steps, "warmup": warmup, "loss": float(loss),
           "model_tflops": tf, "model_tc_frac_of_sustained_peak": tf / pk["tc"], "peak_source": pk["src"],
           "mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}
    print(json.dumps(out), flush=True)
    del stepper, dp, opt, model
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()


def c2():
    import omr_a2s_multimodal_transformer_b200 as pkg

    w2i, i2w = synth.load_vocab()
    m = pkg.Transformer(195, 808, 1268, w2i, i2w)
    m.load_state_dict(synth.synth_state_dict(m.state_dict(), seed=0))
    b, t = 16, 512

    def batch():
        xi, xli, xa, xla, y_in, y_out = bench.make_batch(b, w2i, seed=200, t_len=t)
        return [xa.pin_memory(), xla, y_in, y_out]

    def build(model, dp, opt):
        def step(bt):
            x, xl, y_in, y_out = bt
            dp.zero_grad()
            y_in = model.apply_teacher_forcing(y_in)
            loss = model.decoder.loss(tgt=y_in, memory=model.encode(x), memory_len=xl, targets=y_out)
            loss.backward()
            dp.sync_gradients()
            opt.step()
            return loss

        return step

    d, s, L, ff, v = 256, 1313, 8, 256, 6997
    dec = L * (8 * t * d * d + 4 * t * t * d + 4 * t * d * d + 4 * s * d * d + 4 * t * s * d + 4 * t * d * ff) + 2 * t * d * v
    run("C2 audio-only train step (195x808 spectrograms, batch 16, T=512, bf16)", m, batch, build, 3.0 * (19.551e9 + dec), b)


def c5(t=1024):
    import omr_a2s_multimodal_transformer_b200 as pkg
    from omr_a2s_multimodal_transformer_b200.decoder import Decoder
    from omr_a2s_multimodal_transformer_b200.encoder import Encoder
    from omr_a2s_multimodal_transformer_b200.model import PositionalEncoding2D

    w2i, i2w = synth.load_vocab()
    D, H, L, MAXLEN = 512, 8, 8, 2536
    m = pkg.MultimodalTransformer(128, 1024, 195, 808, MAXLEN, w2i, i2w, mixer_type="concat")
    m.image_encoder = Encoder(1, out_channels=D)
    m.audio_encoder = Encoder(1, out_channels=D)
    m.image_pos_2d = PositionalEncoding2D(D, 8, 128)
    m.audio_pos_2d = PositionalEncoding2D(D, 13, 101)
    m.decoder = Decoder(len(w2i), MAXLEN, len(w2i), embedding_dim=D, ff_dim=D, nhead=H, num_transformer_layers=L,
                        padding_idx=m.padding_idx)
    m.load_state_dict(synth.synth_state_dict(m.state_dict(), seed=0))
    b = 32

    def batch():
        return list(bench.make_batch(b, w2i, seed=300, t_len=t))

    def build(model, dp, opt):
        def step(bt):
            xi, xli, xa, xla, y_in, y_out = bt
            dp.zero_grad()
            y_in = model.apply_teacher_forcing(y_in)
            mem, xl = model._memory(xi, xa, xli, xla, "both")
            loss = model.decoder.loss(tgt=y_in, memory=mem, memory_len=xl, targets=y_out)
            loss.backward()
            dp.sync_gradients()
            opt.step()
            return loss

        return step

    d, s, ff, v = D, 2337, D, 6997
    dec = L * (8 * t * d * d + 4 * t * t * d + 4 * t * d * d + 4 * s * d * d + 4 * t * s * d + 4 * t * d * ff) + 2 * t * d * v
    enc = 16.110e9 + 19.551e9  # + the wider last pointwise convolutions (128 -> 512), < 1 % of the encoders
    run(f"C5 scaled-up multimodal train step (d_model 512, 8 heads, 8 layers, max_len 2536; batch 32, T={t}, S=2337, bf16)",
        m, batch, build, 3.0 * (enc + dec), b)


if __name__ == "__main__":
    which = sys.argv[1:] or ["c2", "c5"]
    for w in which:
        t0 = time.time()
        try:
            {"c2": c2, "c5": c5, "c5long": lambda: c5(2535)}[w]()
        except Exception as e:  # report and go on to the next configuration
            import traceback

            traceback.print_exc()
            print(json.dumps({"config": w, "error": repr(e)[:400]}), flush=True)
        print(f"# {w}: {time.time() - t0:.1f} s", file=sys.stderr, flush=True)
