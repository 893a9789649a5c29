"""-m gpu: the CUDA path against the COMMITTED golden vectors (tests/golden/*.pt = outputs of the REAL reference,
generated in the build container by oracle/make_golden.py): fp32 logits, loss, per-tensor gradient norms and the
reference's own greedy token streams.  No oracle code computes the expected values here."""
import os

import pytest
import torch

from oracle import synth
from tests.helpers import build_multimodal, build_unimodal, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    return torch.load(os.path.join(GOLDEN, f"{name}.pt"), weights_only=False)


def grads_vs_summary(m, summary):
    """Model gradients against a fixture's per-tensor summary {key: (L2 norm, projection on a key-seeded random direction)}
    (oracle/make_golden.py:grad_summary).  Returns
      norm_rel : sqrt(sum (|g_k| - nrm_k)^2 / sum nrm_k^2)                     -- magnitudes
      proj_rel : sqrt(sum (<g_k, r_k> - prj_k)^2 / sum nrm_k^2)                 -- an unbiased estimate of the global L2-relative
                 gradient error (E[<d, r>^2] = |d|^2 for r ~ N(0, I)): unlike the norms it sees signs, transposes, permutations
      worst    : (key, |<g_k, r_k> - prj_k| / nrm_k) over the tensors that carry at least 1e-3 of the global norm"""
    from oracle.make_golden import proj_vec

    grads = dict(m.named_parameters())
    tot = sum(n ** 2 for n, _ in summary.values()) ** 0.5
    num_n = num_p = 0.0
    worst = ("", 0.0)
    for k, (nrm, prj) in summary.items():
        g = grads[k].grad.detach().double().cpu().reshape(-1)
        num_n += (float(g.norm()) - nrm) ** 2
        dp = float((g * proj_vec(k, g.numel())).sum()) - prj
        num_p += dp ** 2
        if nrm > 1e-3 * tot and abs(dp) / nrm > worst[1]:
            worst = (k, abs(dp) / nrm)
    return dict(norm_rel=num_n ** 0.5 / tot, proj_rel=num_p ** 0.5 / tot, worst=worst)


@pytest.mark.parametrize("window", [-1, 5])
def test_unimodal_fp32_matches_reference_vectors(window):
    fix = load(f"uni_w{window}")
    m, sd, w2i = build_unimodal(window=window, dtype=torch.float32)
    x, xl, y_in, y_out = synth.synth_unimodal_batch(3, 64, 128, [20, 12, 7], w2i)
    with torch.no_grad():
        logits = m(x.to(DEV), xl.to(DEV), y_in.to(DEV))
    assert rel_err(logits, fix["logits_fp64"]) < 1e-4
    m.zero_grad(set_to_none=True)
    loss = m.decoder.loss(tgt=y_in.to(DEV), memory=m.encode(x.to(DEV)), memory_len=xl.to(DEV), targets=y_out.to(DEV))
    loss.backward()
    assert abs(float(loss) - fix["loss_ref_fp32"]) < 1e-4 * max(1.0, abs(fix["loss_ref_fp32"]))
    rep = grads_vs_summary(m, fix["grad_summary_ref_fp32"])  # the real reference's own fp32 run
    assert rep["norm_rel"] < 1e-3 and rep["proj_rel"] < 2.5e-3 and rep["worst"][1] < 5e-2, rep
    toks, vals, lens = m.greedy_decode_batch(x.to(DEV), max_steps=24)
    seqs, _ = m._decoder_runner().to_lists(toks, vals, lens)
    for i, ref_tokens in enumerate(fix["greedy"]):
        assert seqs[i] == ref_tokens, (i, seqs[i], ref_tokens)


@pytest.mark.parametrize("mixer", ["concat", "attn_img", "attn_audio", "attn_both"])
def test_multimodal_fp32_and_bf16_match_reference_vectors(mixer):
    fix = load(f"mm_{mixer}")
    m, sd, w2i = build_multimodal(mixer=mixer, dtype=torch.float32)
    xi, xli, xa, xla, y_in, y_out = synth.synth_multimodal_batch(3, (64, 128), (48, 96), [20, 12, 7], w2i)
    args = (xi.to(DEV), xli.to(DEV), xa.to(DEV), xla.to(DEV), y_in.to(DEV))
    with torch.no_grad():
        logits = m(*args)
    assert rel_err(logits, fix["logits_ref_fp32"]) < 1e-4
    m.set_compute_dtype(torch.bfloat16)
    with torch.no_grad():
        lb = m(*args)
    # bf16: the north-star 1e-2, or 1.5x the reference's own autocast-bf16 error on these inputs if that is larger
    assert rel_err(lb.float(), fix["logits_ref_fp32"]) < max(1e-2, 1.5 * fix["noise"]["bf16_logits"])
    if "greedy" in fix:
        m.set_compute_dtype(torch.float32)
        toks, vals, lens = m.greedy_decode_batch(xi.to(DEV), xa.to(DEV), max_steps=24)
        seqs, _ = m._decoder_runner().to_lists(toks, vals, lens)
        for i, ref_tokens in enumerate(fix["greedy"]):
            assert seqs[i] == ref_tokens, (i, seqs[i], ref_tokens)


# ---- BASELINE configs at their real per-sample shapes (grandstaff vocabulary, max_len 1268) ---------------------------
def _full_size_model(kind, seed):
    import omr_a2s_multimodal_transformer_b200 as pkg

    w2i, i2w = synth.load_vocab()
    if kind == "c1":
        m = pkg.Transformer(128, 1024, 1268, w2i, i2w)
    elif kind == "c2":
        m = pkg.Transformer(195, 808, 1268, w2i, i2w)
    else:
        m = pkg.MultimodalTransformer(128, 1024, 195, 808, 1268, w2i, i2w, mixer_type="concat")
    m.load_state_dict(synth.synth_state_dict(m.state_dict(), seed=seed))
    m = m.to(DEV).eval()
    m.set_compute_dtype(torch.float32)
    return m, w2i


def _check_full_size(fix, m, fwd, loss_fn, greedy_fn):
    sv, st = fix["logits_stride"]
    with torch.no_grad():
        logits = fwd()
    assert tuple(logits.shape) == fix["logits_shape"]
    assert rel_err(logits[:, ::sv, ::st], fix["logits_fp64"]) < 1e-4
    assert rel_err(logits[:, ::sv, ::st], fix["logits_ref_fp32"]) < 1e-4
    m.zero_grad(set_to_none=True)
    loss = loss_fn()
    loss.backward()
    assert abs(float(loss) - fix["loss_ref_fp32"]) < 1e-4 * max(1.0, abs(fix["loss_ref_fp32"]))
    # per-tensor gradient norms AND projections against the fp64 truth; the reference's own fp32 run is ref_floor away from
    # it.  The projection figure estimates the global relative error from one random direction per tensor (a few dominant
    # tensors => it scatters by ~2x), hence the 2.5x; a sign / transpose / permutation bug moves it to O(1).
    ref_floor = fix["noise"]["fp32_grad"][0]
    rep = grads_vs_summary(m, fix["grad_summary_fp64"])
    assert rep["norm_rel"] < max(1e-3, 4 * ref_floor), rep
    assert rep["proj_rel"] < 2.5 * max(1e-3, 4 * ref_floor) and rep["worst"][1] < 5e-2, rep
    seqs = greedy_fn()
    for i, ref_tokens in enumerate(fix["greedy"]):
        assert seqs[i] == ref_tokens, (i, seqs[i], ref_tokens)
    m.set_compute_dtype(torch.bfloat16)
    with torch.no_grad():
        lb = fwd()
    assert rel_err(lb.float()[:, ::sv, ::st], fix["logits_fp64"]) < max(1e-2, 1.5 * fix["noise"]["bf16_logits"])
    # bf16 gradients at the full-size shapes: global relative error (projection estimate) and magnitudes against the
    # north-star 1e-2 or 1.5x the REAL reference's own autocast-bf16 gradient error on these inputs, whichever is larger
    m.zero_grad(set_to_none=True)
    loss_b = loss_fn()
    loss_b.backward()
    assert abs(float(loss_b) - fix["loss_fp64"]) < max(1e-2, 1.5 * fix["noise"]["bf16_loss"]) * max(1.0, abs(fix["loss_fp64"]))
    floor_b = max(1e-2, 1.5 * fix["noise"]["bf16_grad"][0])
    rep = grads_vs_summary(m, fix["grad_summary_fp64"])
    assert rep["norm_rel"] < floor_b and rep["proj_rel"] < 2.5 * floor_b and rep["worst"][1] < 3.0, (rep, floor_b)


@pytest.mark.parametrize("name,seed,hw,lens,frames,pad", [
    ("c1_image_only", 0, (128, 1024), [257, 200, 128, 64], [1024, 1024, 896, 768], 1.0),
    ("c2_audio_only", 5, (195, 808), [300, 129], [1313, 900], 0.0),
])
def test_full_size_unimodal_matches_reference_vectors(name, seed, hw, lens, frames, pad):
    """BASELINE config 1 (image-only, batch 4) and config 2's shapes (audio-only 195x808 spectrograms)"""
    fix = load(name)
    m, w2i = _full_size_model(name[:2], seed)
    x, xl, y_in, y_out = synth.synth_unimodal_batch(len(lens), hw[0], hw[1], lens, w2i, pad_value=pad, frame_lens=frames)
    x, xl, y_in, y_out = x.to(DEV), xl.to(DEV), y_in.to(DEV), y_out.to(DEV)
    steps = 48 if name.startswith("c1") else 32

    def greedy():
        toks, vals, ln = m.greedy_decode_batch(x[:1], max_steps=steps)
        return m._decoder_runner().to_lists(toks, vals, ln)[0]

    _check_full_size(fix, m, lambda: m(x, xl, y_in),
                     lambda: m.decoder.loss(tgt=y_in, memory=m.encode(x), memory_len=xl, targets=y_out), greedy)


def test_full_size_multimodal_matches_reference_vectors():
    """BASELINE config 3's shapes: image 1x128x1024 + audio 1x195x808 -> fused memory of 2337 positions"""
    fix = load("c3_multimodal")
    m, w2i = _full_size_model("c3", 6)
    batch = synth.synth_multimodal_batch(2, (128, 1024), (195, 808), [300, 129], w2i, img_frame_lens=[1024, 700],
                                         aud_frame_lens=[800, 1313])
    xi, xli, xa, xla, y_in, y_out = (t.to(DEV) for t in batch)

    def loss_fn():
        mem, xl = m._memory(xi, xa, xli, xla, "both")
        return m.decoder.loss(tgt=y_in, memory=mem, memory_len=xl, targets=y_out)

    def greedy():
        toks, vals, ln = m.greedy_decode_batch(xi[:1], xa[:1], max_steps=32)
        return m._decoder_runner().to_lists(toks, vals, ln)[0]

    _check_full_size(fix, m, lambda: m(xi, xli, xa, xla, y_in), loss_fn, greedy)


# ---- long greedy decodes: every sequence of the batch, to <eos> / max_seq_len ---------------------
TIE_TOL = 2e-5  # a top-2 logit margin below this (|logit| ~ 3: ~7e-6 relative, a few fp32 ulps of the accumulations) is a tie


def _assert_identical_streams(seqs, fix):
    """token-for-token identity with the REAL reference's batch-1 loop.  The fixture carries the reference's own top-2 logit
    margin at every step (oracle/make_golden_greedy.py): a stream may leave the reference's ONLY at a step where the
    reference itself sits within TIE_TOL of a tie (its fp32 CPU run and any other correctly rounded fp32 evaluation may
    then pick either token; one such step exists in 6 x ~1000 steps: sample 2 of config 1, margin 2.1e-6) -- everything up
    to that step must be identical, and the divergence is reported as a warning.  Any other mismatch fails with the margin."""
    import warnings

    assert len(seqs) >= len(fix["greedy"])
    for i, ref in enumerate(fix["greedy"]):
        got = seqs[i]
        if got == ref:
            continue
        t = next((k for k in range(min(len(got), len(ref))) if got[k] != ref[k]), min(len(got), len(ref)))
        mg = fix["margins"][i]
        if t < len(mg) and float(mg[t]) < TIE_TOL and len(got) == len(ref):
            warnings.warn(f"sample {i}: identical for {t} steps, then a reference near-tie (top-2 margin {float(mg[t]):.2e} < {TIE_TOL})")
            continue
        raise AssertionError(f"sample {i}: diverges from the reference at step {t} of {len(ref)} (got {len(got)} tokens); reference top-2 "
                             f"margin there {float(mg[t]) if t < len(mg) else None:.3e}, min margin of the stream {float(mg.min()):.3e}, "
                             f"|logit| max {fix['logit_scale'][i]:.2f}")


def test_greedy_long_c1_all_four_samples_to_max_len():
    """BASELINE config 1: all 4 image-only samples decoded as ONE batch to max_seq_len = 1268 (random-init models never emit
    <eos>) == the reference's own validation_step loop run on each sample alone (fp32)"""
    fix = load("greedy_long_c1")
    m, w2i = _full_size_model("c1", 0)
    x = synth.synth_unimodal_batch(4, 128, 1024, [257, 200, 128, 64], w2i, frame_lens=[1024, 1024, 896, 768])[0].to(DEV)
    toks, vals, ln = m.greedy_decode_batch(x, max_steps=fix["steps"])
    seqs = m._decoder_runner().to_lists(toks, vals, ln)[0]
    assert [len(s) for s in seqs] == [len(r) for r in fix["greedy"]]
    _assert_identical_streams(seqs, fix)


def test_greedy_long_c3_both_samples_640_steps():
    """BASELINE config 3's shapes (fused memory of 2337 positions): both samples of the fixture batch, 640 steps (fp32)"""
    fix = load("greedy_long_c3")
    m, w2i = _full_size_model("c3", 6)
    batch = synth.synth_multimodal_batch(2, (128, 1024), (195, 808), [300, 129], w2i, img_frame_lens=[1024, 700],
                                         aud_frame_lens=[800, 1313])
    xi, xa = batch[0].to(DEV), batch[2].to(DEV)
    toks, vals, ln = m.greedy_decode_batch(xi, xa, max_steps=fix["steps"])
    seqs = m._decoder_runner().to_lists(toks, vals, ln)[0]
    _assert_identical_streams(seqs, fix)
