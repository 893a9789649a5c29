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
    grads = dict(m.named_parameters())
    num = den = 0.0
    for k, (nrm, _prj) in fix["grad_summary_ref_fp32"].items():
        got = float(grads[k].grad.double().norm())
        num += (got - nrm) ** 2
        den += nrm ** 2
    assert (num / den) ** 0.5 < 1e-3  # per-tensor gradient norms of the real reference's fp32 run
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
