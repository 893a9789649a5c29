"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.pt by running the REAL reference
(/root/reference, via oracle/shim.py) on the synthetic weights/inputs of oracle/synth.py.

Run in the build container:  python -m oracle.make_golden
Each fixture holds, for one configuration: fp64 reference logits (possibly strided), loss, a per-tensor
gradient summary (L2 norm + projection on a key-seeded random direction), the reference's greedy token
streams, and the reference's OWN bf16-autocast error against its fp64 run (the noise floor that the
bf16 tolerances of the parity tests are expressed against).
"""
from __future__ import annotations

import os
import sys
import zlib

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import shim, synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def proj_vec(key: str, n: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(zlib.crc32(key.encode()) & 0x7FFFFFFF)
    return torch.randn(n, generator=g, dtype=torch.float64)


def grad_summary(named_grads):
    out = {}
    for k, g in named_grads:
        if g is None:
            continue
        g = g.detach().double().reshape(-1)
        out[k] = (float(g.norm()), float((g * proj_vec(k, g.numel())).sum()))
    return out


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-300))


def run_case(name, build, batch_fn, fwd, oracle_fwd, greedy_inputs=None, logits_stride=(1, 1), greedy_steps=None):
    """fwd(model, batch) runs the REAL reference; oracle_fwd(sd, batch, dtype) runs oracle/restate.py.

    torch's CPU SDPA mis-handles an fp32 float mask next to fp64 queries (the reference hard-codes fp32
    masks, decoder.py:185,253), so the real reference cannot serve as its own fp64 ground truth; the
    fp64 ground truth is the restatement, which this script first checks against the real reference in
    fp32 (max-norm relative difference must be at rounding level)."""
    from oracle import restate

    ref = shim.load_reference()
    torch.manual_seed(0)
    model, w2i = build(ref)
    model.eval()
    sd = synth.synth_state_dict(model.state_dict(), seed=build.seed)
    model.load_state_dict(sd)
    batch = batch_fn(w2i)
    y_out = batch[-1]
    # the real reference, fp32
    model.zero_grad()
    logits32 = fwd(model, batch)
    loss32 = model.compute_loss(logits32, y_out)
    loss32.backward()
    g32 = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
    # restatement: fp32 pin check, fp64 ground truth
    with torch.no_grad():
        pin = rel(oracle_fwd(sd, batch, torch.float32), logits32)
    assert pin < 5e-6, f"{name}: restatement deviates from the real reference ({pin})"
    sdg = {k: (v.double().requires_grad_(True) if torch.is_floating_point(v) and not k.endswith(".pe") else v) for k, v in sd.items()}
    logits64 = oracle_fwd(sdg, batch, torch.float64)
    loss64 = restate.ce_loss(logits64, y_out)
    loss64.backward()
    g64 = {k: sdg[k].grad for k in g32}
    # the reference's own bf16 autocast run
    model.zero_grad()
    with torch.autocast("cpu", dtype=torch.bfloat16):
        logits16 = fwd(model, batch)
        loss16 = model.compute_loss(logits16.float(), y_out)
    loss16.backward()
    g16 = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}

    def gerr(ga):
        num = sum(float((ga[k].double() - g64[k]).pow(2).sum()) for k in g64)
        den = sum(float(g64[k].pow(2).sum()) for k in g64)
        dot = sum(float((ga[k].double() * g64[k]).sum()) for k in g64)
        na = sum(float(ga[k].double().pow(2).sum()) for k in g64)
        return (num / den) ** 0.5, dot / (na * den) ** 0.5

    fix = {
        "name": name,
        "logits_stride": logits_stride,
        "logits_ref_fp32": logits32.detach()[:, :: logits_stride[0], :: logits_stride[1]].float().clone(),
        "logits_fp64": logits64.detach()[:, :: logits_stride[0], :: logits_stride[1]].clone(),
        "logits_shape": tuple(logits64.shape),
        "loss_ref_fp32": float(loss32.detach()),
        "loss_fp64": float(loss64.detach()),
        "grad_summary_ref_fp32": grad_summary(g32.items()),
        "grad_summary_fp64": grad_summary(g64.items()),
        "pin_restate_vs_reference_fp32": pin,
        "noise": {
            "fp32_logits": rel(logits32, logits64), "fp32_loss": abs(float(loss32) - float(loss64)),
            "fp32_grad": gerr(g32), "bf16_logits": rel(logits16.float(), logits64),
            "bf16_loss": abs(float(loss16) - float(loss64)), "bf16_grad": gerr(g16),
        },
    }
    # greedy: the reference's own validation_step loop (fp32)
    if greedy_inputs is not None:
        seqs = []
        saved = model.max_seq_len
        if greedy_steps is not None:
            model.max_seq_len = greedy_steps
        for gb in greedy_inputs(w2i, batch):
            model.Y, model.YHat = [], []
            model.validation_step(gb, 0)
            seqs.append([w2i[t] for t in model.YHat[0]])
        model.max_seq_len = saved
        fix["greedy"] = seqs
    os.makedirs(OUT, exist_ok=True)
    torch.save(fix, os.path.join(OUT, f"{name}.pt"))
    n = fix["noise"]
    print(f"{name}: loss {fix['loss_fp64']:.6f} pin {pin:.1e} | ref noise fp32 logits {n['fp32_logits']:.2e} "
          f"grad {n['fp32_grad'][0]:.2e} | bf16 logits {n['bf16_logits']:.2e} loss {n['bf16_loss']:.2e} grad {n['bf16_grad'][0]:.2e} cos {n['bf16_grad'][1]:.5f}")


def builder(kind, seed, **kw):
    def build(ref):
        w2i, i2w = synth.tiny_vocab(97) if kw.get("vocab", "tiny") == "tiny" else synth.load_vocab()
        if kind == "uni":
            m = ref.Transformer(kw["hw"][0], kw["hw"][1], kw["max_len"], w2i, i2w, attn_window=kw.get("window", -1))
        else:
            m = ref.MultimodalTransformer(kw["img"][0], kw["img"][1], kw["aud"][0], kw["aud"][1], kw["max_len"], w2i, i2w,
                                          mixer_type=kw.get("mixer", "concat"), attn_window=kw.get("window", -1))
        return m, w2i

    build.seed = seed
    return build


def main():
    from oracle import restate

    if len(sys.argv) > 1 and sys.argv[1] == "full-size":
        return main_full_size()
    lens = [20, 12, 7]
    for window in (-1, 5):
        run_case(
            f"uni_w{window}", builder("uni", 4, hw=(64, 128), max_len=40, window=window),
            lambda w2i: synth.synth_unimodal_batch(3, 64, 128, lens, w2i),
            lambda m, b: m(b[0], b[1], b[2]),
            lambda sd, b, dt, window=window: restate.unimodal_forward(sd, b[0], b[1], b[2], attn_window=window, dtype=dt),
            greedy_inputs=lambda w2i, b: [(b[0][i:i + 1], torch.tensor([[w2i["<sos>"], 3, 4, w2i["<eos>"]]])) for i in range(3)],
            greedy_steps=24,
        )
    for mixer in ("concat", "attn_img", "attn_audio", "attn_both"):
        run_case(
            f"mm_{mixer}", builder("mm", 3, img=(64, 128), aud=(48, 96), max_len=40, mixer=mixer),
            lambda w2i: synth.synth_multimodal_batch(3, (64, 128), (48, 96), lens, w2i),
            lambda m, b: m(b[0], b[1], b[2], b[3], b[4]),
            lambda sd, b, dt, mixer=mixer: restate.multimodal_forward(sd, b[0], b[1], b[2], b[3], b[4], mixer_type=mixer, dtype=dt),
            greedy_inputs=(lambda w2i, b: [(b[0][i:i + 1], b[2][i:i + 1], torch.tensor([[w2i["<sos>"], 3, 4, w2i["<eos>"]]]))
                                           for i in range(2)]) if mixer == "concat" else None,
            greedy_steps=24,
        )
    # C1 of BASELINE.json: image-only, 1x128x1024, batch 4, real vocabulary, fp32 CPU (logits strided to keep the fixture small)
    run_case(
        "c1_image_only", builder("uni", 0, hw=(128, 1024), max_len=1268, vocab="real"),
        lambda w2i: synth.synth_unimodal_batch(4, 128, 1024, [257, 200, 128, 64], w2i, frame_lens=[1024, 1024, 896, 768]),
        lambda m, b: m(b[0], b[1], b[2]),
        lambda sd, b, dt: restate.unimodal_forward(sd, b[0], b[1], b[2], dtype=dt),
        greedy_inputs=lambda w2i, b: [(b[0][:1], torch.tensor([[w2i["<sos>"], 3, 4, w2i["<eos>"]]]))],
        logits_stride=(97, 8), greedy_steps=48,
    )


def main_full_size():
    """BASELINE configs 2 and 3 at their real per-sample shapes (195x808 spectrograms, 128x1024 images, the grandstaff
    vocabulary and max length), batch 2 so that the fp64 run fits the build container; logits strided."""
    from oracle import restate

    run_case(
        "c2_audio_only", builder("uni", 5, hw=(195, 808), max_len=1268, vocab="real"),
        lambda w2i: synth.synth_unimodal_batch(2, 195, 808, [300, 129], w2i, pad_value=0.0, frame_lens=[1313, 900]),
        lambda m, b: m(b[0], b[1], b[2]),
        lambda sd, b, dt: restate.unimodal_forward(sd, b[0], b[1], b[2], dtype=dt),
        greedy_inputs=lambda w2i, b: [(b[0][:1], torch.tensor([[w2i["<sos>"], 3, 4, w2i["<eos>"]]]))],
        logits_stride=(97, 8), greedy_steps=32,
    )
    run_case(
        "c3_multimodal", builder("mm", 6, img=(128, 1024), aud=(195, 808), max_len=1268, mixer="concat", vocab="real"),
        lambda w2i: synth.synth_multimodal_batch(2, (128, 1024), (195, 808), [300, 129], w2i, img_frame_lens=[1024, 700],
                                                 aud_frame_lens=[800, 1313]),
        lambda m, b: m(b[0], b[1], b[2], b[3], b[4]),
        lambda sd, b, dt: restate.multimodal_forward(sd, b[0], b[1], b[2], b[3], b[4], mixer_type="concat", dtype=dt),
        greedy_inputs=lambda w2i, b: [(b[0][:1], b[2][:1], torch.tensor([[w2i["<sos>"], 3, 4, w2i["<eos>"]]]))],
        logits_stride=(97, 8), greedy_steps=32,
    )


if __name__ == "__main__":
    main()
