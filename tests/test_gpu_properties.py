"""-m gpu: size-independent properties of
the CUDA path at BASELINE.json's real per-sample shapes (image 1x128x1024, audio 1x195x808, grandstaff vocabulary), where
the CPU oracle is too slow to serve as the checker:
  * batch independence  -- a sample's logits do not depend on its batch mates (InstanceNorm per sample, LayerNorm per
    token, attention per sequence: SURVEY.md section 8e);
  * causality           -- logits at positions < t0 do not depend on the tokens at positions >= t0;
  * key-padding algebra -- under the concat mixer's BOOL mask the padded memory frames are truly excluded, under the
    unimodal INT-length mask they are not (the float mask is additive, +1.0: SURVEY.md section 8 a9) -- both as the reference;
  * batched greedy decoding == one sample at a time."""
import pytest
import torch

from oracle import synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _full_size(kind, dtype):
    import omr_a2s_multimodal_transformer_b200 as pkg

    w2i, i2w = synth.load_vocab()
    if kind == "mm":
        m = pkg.MultimodalTransformer(128, 1024, 195, 808, 1268, w2i, i2w, mixer_type="concat")
    else:
        m = pkg.Transformer(128, 1024, 1268, w2i, i2w)
    m.load_state_dict(synth.synth_state_dict(m.state_dict(), seed=7))
    m = m.to(DEV).eval()
    m.set_compute_dtype(dtype)
    return m, w2i


def _rel(a, b):
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-30))


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 2e-2)])
def test_batch_independence_at_c3_shapes(dtype, tol):
    m, w2i = _full_size("mm", dtype)
    batch = synth.synth_multimodal_batch(6, (128, 1024), (195, 808), [300, 129, 64, 257, 33, 200], w2i)
    xi, xli, xa, xla, y_in, y_out = (t.to(DEV) for t in batch)
    with torch.no_grad():
        full = m(xi, xli, xa, xla, y_in)
        for k in (1, 4):
            one = m(xi[k:k + 1], xli[k:k + 1], xa[k:k + 1], xla[k:k + 1], y_in[k:k + 1])
            assert _rel(full[k:k + 1], one) < tol, (k, _rel(full[k:k + 1], one))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_causality_at_c3_shapes(dtype):
    m, w2i = _full_size("mm", dtype)
    batch = synth.synth_multimodal_batch(3, (128, 1024), (195, 808), [400, 400, 400], w2i)
    xi, xli, xa, xla, y_in, y_out = (t.to(DEV) for t in batch)
    t0 = 173
    y2 = y_in.clone()
    y2[:, t0:] = torch.randint(1, 6000, y2[:, t0:].shape, device=DEV)
    with torch.no_grad():
        a = m(xi, xli, xa, xla, y_in)
        b = m(xi, xli, xa, xla, y2)
    # same shapes, same launch geometry: the visible prefix is computed by the same instructions on the same data
    assert torch.equal(a[:, :, :t0], b[:, :, :t0])
    assert not torch.equal(a[:, :, t0:], b[:, :, t0:])


def test_key_padding_algebra_matches_the_reference_semantics():
    m, w2i = _full_size("mm", torch.float32)
    xi, xli, xa, xla, y_in, _ = (t.to(DEV) for t in synth.synth_multimodal_batch(
        2, (128, 1024), (195, 808), [90, 40], w2i, img_frame_lens=[1024, 600], aud_frame_lens=[700, 1313]))
    with torch.no_grad():
        mem, bias = m._memory(xi, xa, xli, xla, "both")
        a = m.decoder(tgt=y_in, memory=mem, memory_len=bias)
        mem2 = mem.clone()
        mem2[1, 600:1024] += 3.0          # padded image frames of sample 1
        mem2[0, 1024 + 700:] -= 2.0       # padded audio frames of sample 0
        b = m.decoder(tgt=y_in, memory=mem2, memory_len=bias)
    assert torch.equal(a, b)  # bool mask of the concat mixer: padded frames are excluded (bias = -inf)
    u, _ = _full_size("uni", torch.float32)
    x = xi
    xl = torch.tensor([1024, 600], dtype=torch.int32, device=DEV)
    with torch.no_grad():
        memu = u.encode(x)
        c = u.decoder(tgt=y_in, memory=memu, memory_len=xl)
        memu2 = memu.clone()
        memu2[1, 600:] += 3.0
        d = u.decoder(tgt=y_in, memory=memu2, memory_len=xl)
    assert torch.equal(c[0], d[0])      # sample 0 has no padded frame
    assert not torch.equal(c[1], d[1])  # int lengths -> additive float mask (+1.0): padded frames still take part


def test_batched_greedy_equals_single_sample_greedy_fp32():
    m, w2i = _full_size("uni", torch.float32)
    x = torch.rand(3, 1, 128, 1024, generator=torch.Generator().manual_seed(3)).to(DEV)
    toks, vals, lens = m.greedy_decode_batch(x, max_steps=40)
    seqs, _ = m._decoder_runner().to_lists(toks, vals, lens)
    for k in range(3):
        t1, v1, l1 = m.greedy_decode_batch(x[k:k + 1], max_steps=40)
        s1, _ = m._decoder_runner().to_lists(t1, v1, l1)
        assert s1[0] == seqs[k], (k, s1[0], seqs[k])
