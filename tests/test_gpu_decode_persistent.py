"""The persistent decode kernel (csrc/decode_persistent.cu) beyond the small shapes of test_gpu_model.py (-m gpu):
the weight layouts its bf16 tensor-core projections read, and the bf16 kernel against the per-kernel decode path and the
fp32 kernel at the full C4 memory length (S = 2337: 147 sixteen-key tiles, nine per warp, ragged last tile), with the real
vocabulary size (classifier slots with a ragged tail) and with a key-padding bias on the memory.  The reference decodes
one unpadded sample at a time (src/transformer/model.py:170-199, 592-617); the fp32 token identity against its own loop is
pinned by tests/test_gpu_golden.py."""
import pytest
import torch

from oracle import synth
from tests.helpers import build_multimodal

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_decode_weight_layouts_match_the_documented_index_formulas():
    """params.WeightCache kinds matKS4 / matDecA / matDecKS against include/omr_b200.h (omr_decode_layer):
    column slices [4][N][K/4]; A-fragment order [rows/16][K/32][2][8][4][8] holding W[16 tile + 8 hf + g][32 kb + 8 t + e]."""
    from omr_a2s_multimodal_transformer_b200.params import WeightCache

    g = torch.Generator().manual_seed(7)
    for n, k in ((256, 256), (768, 256), (97, 256)):
        w = torch.randn(n, k, generator=g).to(DEV)
        wb = w.to(torch.bfloat16)
        if n % 16 == 0:
            ks = WeightCache._pack(w, "matKS4", torch.float32).view(4, n, k // 4)
            for r in range(4):
                assert torch.equal(ks[r], w[:, r * (k // 4):(r + 1) * (k // 4)])
            dks = WeightCache._pack(w, "matDecKS", torch.bfloat16).view(4, n // 16, (k // 4) // 32, 2, 8, 4, 8).cpu()
            ref = wb.cpu()
            for (r, tile, kb, hf, gg, t, e) in ((0, 0, 0, 0, 0, 0, 0), (3, n // 16 - 1, 1, 1, 7, 3, 7), (1, 2, 0, 1, 3, 2, 5), (2, 5, 1, 0, 6, 1, 2)):
                assert dks[r, tile, kb, hf, gg, t, e] == ref[16 * tile + 8 * hf + gg, r * (k // 4) + 32 * kb + 8 * t + e]
        npad = (n + 31) // 32 * 32
        da = WeightCache._pack(w, "matDecA", torch.bfloat16)
        assert da.shape == (npad, k)
        da = da.view(npad // 16, k // 32, 2, 8, 4, 8).cpu()
        full = torch.zeros(npad, k, dtype=torch.bfloat16)
        full[:n] = wb.cpu()
        # every element, through the inverse permutation
        back = da.permute(0, 2, 3, 1, 4, 5).reshape(npad, k)
        assert torch.equal(back, full)


def _decode_all_paths(monkeypatch, build_kw, batch, lens, steps, bias_tail):
    runs = {}
    for name, dtype, mode in (("bf16_persistent", torch.bfloat16, "persistent"), ("bf16_graph", torch.bfloat16, "graph"),
                              ("fp32_persistent", torch.float32, "persistent")):
        monkeypatch.setenv("OMR_DECODE_MODE", mode)
        m, sd, w2i = build_multimodal(dtype=dtype, **build_kw)
        xi, _, xa, _, _, _ = synth.synth_multimodal_batch(batch, build_kw["img"], build_kw["aud"], lens, w2i)
        with torch.no_grad():
            mem, _ = m._memory(xi.to(DEV), xa.to(DEV), None, None, "both")
            bias = None
            if bias_tail:
                # sample b ignores the last bias_tail[b] memory positions (a whole tile and a ragged part of the next)
                s = mem.shape[1]
                bias = torch.zeros(batch, s, dtype=torch.float32, device=DEV)
                for b, tail in enumerate(bias_tail):
                    if tail:
                        bias[b, s - tail:] = float("-inf")
            toks, vals, _ = m._decoder_runner().decode(mem, w2i["<sos>"], w2i["<eos>"], 0, max_steps=steps, stop_at_eos=False,
                                                       mem_bias=bias)
        runs[name] = (toks.cpu(), vals.cpu())
    return runs


def _compare(runs, batch, steps, tol=3e-2, min_agree_graph=None):
    ref_t, ref_v = runs["bf16_persistent"]
    assert ref_t.shape == (batch, steps)
    for other in ("bf16_graph", "fp32_persistent"):
        t, v = runs[other]
        agree = 0
        for b in range(batch):
            for i in range(steps):
                # top-logit values agree to bf16 accuracy for as long as the token prefixes agree
                assert abs(float(ref_v[b, i]) - float(v[b, i])) <= tol * max(1.0, abs(float(v[b, i]))), (other, b, i)
                agree += 1
                if int(ref_t[b, i]) != int(t[b, i]):
                    break
        if other == "bf16_graph" and min_agree_graph is not None:
            assert agree >= min_agree_graph, (other, agree)


def test_greedy_bf16_persistent_kernel_at_the_full_memory_length_and_vocabulary(monkeypatch):
    """S = 512 + 1825 = 2337 keys, V = 6997 (219 classifier slots of 32 rows over four CTAs, ragged last slot: 6997 % 32 = 21,
    % 4 = 1 -> one bias read directly), 24 steps, batch 2"""
    kw = dict(img=(128, 1024), aud=(195, 808), max_len=24, vocab=6997)
    runs = _decode_all_paths(monkeypatch, kw, 2, [5, 5], 24, None)
    _compare(runs, 2, 24, min_agree_graph=2 * 3)


def test_greedy_persistent_kernel_with_a_key_padding_bias_on_the_memory(monkeypatch):
    """-inf bias on the tail of the memory (an extension: the reference decodes unpadded samples): one sample unmasked, one
    with 40 masked keys (two whole tiles and a ragged part), one with 7; S = 512 + 36 keys"""
    kw = dict(img=(64, 1024), aud=(48, 96), max_len=32)
    runs = _decode_all_paths(monkeypatch, kw, 3, [5, 5, 5], 32, [0, 40, 7])
    _compare(runs, 3, 32, min_agree_graph=3 * 3)
    # the bias matters: the masked runs differ from an unmasked run of the same model
    plain = _decode_all_paths(monkeypatch, kw, 3, [5, 5, 5], 32, None)
    assert not torch.equal(plain["fp32_persistent"][1][1], runs["fp32_persistent"][1][1])
    assert torch.equal(plain["fp32_persistent"][1][0], runs["fp32_persistent"][1][0])


@pytest.mark.parametrize("mode", ["persistent", "graph"])
def test_batched_decode_of_a_ragged_batch_agrees_with_the_teacher_forced_forward(monkeypatch, mode):
    """greedy_decode_batch(xi, xa, xli=..., xla=...) masks the padded memory exactly as forward(xi, xli, xa, xla, y) does
    (decoder.py:128-132 of the reference: the key-padding mask of the fused memory): feeding the decoded prefix back through
    the teacher-forced forward reproduces every decoded token as the argmax of its logits (fp32; near ties excepted)."""
    monkeypatch.setenv("OMR_DECODE_MODE", mode)
    img, aud = (64, 256), (48, 96)
    m, sd, w2i = build_multimodal(dtype=torch.float32, img=img, aud=aud, max_len=24)
    xi, xli, xa, xla, _, _ = synth.synth_multimodal_batch(3, img, aud, [5, 5, 5], w2i)
    assert int(xli.min()) < int(xli.max()) or int(xla.min()) < int(xla.max())  # the batch IS ragged
    xi, xli, xa, xla = xi.to(DEV), xli.to(DEV), xa.to(DEV), xla.to(DEV)
    with torch.no_grad():
        toks, vals, _ = m.greedy_decode_batch(xi, xa, max_steps=16, stop_at_eos=False, xli=xli, xla=xla)
        sos = torch.full((3, 1), w2i["<sos>"], dtype=torch.long, device=DEV)
        y_in = torch.cat([sos, toks[:, :-1]], dim=1)
        logits = m(xi, xli, xa, xla, y_in)  # [B,V,T]
    top2 = logits.float().topk(2, dim=1).values  # [B,2,T]
    for b in range(3):
        for t in range(16):
            margin = float(top2[b, 0, t] - top2[b, 1, t])
            if margin > 1e-4 * max(1.0, abs(float(top2[b, 0, t]))):
                assert int(logits[b, :, t].argmax()) == int(toks[b, t]), (b, t)
            assert abs(float(top2[b, 0, t]) - float(vals[b, t])) <= 1e-4 * max(1.0, abs(float(vals[b, t]))), (b, t)
