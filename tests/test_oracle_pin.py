"""CPU, build container only: pins oracle/restate.py against the REAL reference imported from
/root/reference (skipped where the reference tree is absent, e.g. on the GPU box)."""
import pytest
import torch

from oracle import restate, shim, synth

pytestmark = pytest.mark.skipif(not shim.reference_available(), reason="/root/reference not present")


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-300))


@pytest.fixture(scope="module")
def ref():
    return shim.load_reference()


def test_encoder_and_pe_tables(ref):
    enc = ref.Encoder(1).eval()
    sd = synth.synth_state_dict(enc.state_dict(), seed=5)
    enc.load_state_dict(sd)
    x = torch.rand(2, 1, 50, 77, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        assert rel(restate.encoder_forward(sd, "", x), enc(x)) < 5e-6
    assert torch.equal(restate.pe2d_table(256, 8, 16), ref.PositionalEncoding2D(256, 8, 16).pe)
    assert torch.equal(restate.pe1d_table(40, 256), ref.PositionalEncoding1D(40, 256).pe)


@pytest.mark.parametrize("mixer", ["concat", "attn_both"])
def test_multimodal_forward_and_grads(ref, mixer):
    w2i, i2w = synth.tiny_vocab(97)
    m = ref.MultimodalTransformer(64, 128, 48, 96, 40, w2i, i2w, mixer_type=mixer).eval()
    sd = synth.synth_state_dict(m.state_dict(), seed=3)
    m.load_state_dict(sd)
    xi, xli, xa, xla, y_in, y_out = synth.synth_multimodal_batch(3, (64, 128), (48, 96), [20, 12, 7], w2i)
    out = m(xi, xli, xa, xla, y_in)
    m.compute_loss(out, y_out).backward()
    sdg = {k: (v.clone().requires_grad_(True) if torch.is_floating_point(v) and not k.endswith(".pe") else v) for k, v in sd.items()}
    mine = restate.multimodal_forward(sdg, xi, xli, xa, xla, y_in, mixer_type=mixer)
    restate.ce_loss(mine, y_out).backward()
    assert rel(mine, out) < 5e-6
    num = den = 0.0
    for k, p in m.named_parameters():
        num += float((sdg[k].grad.double() - p.grad.double()).pow(2).sum())
        den += float(p.grad.double().pow(2).sum())
    assert (num / den) ** 0.5 < 1e-4


@pytest.mark.parametrize("modality", ["image", "audio"])
def test_single_modality_memory_and_masks(ref, modality, monkeypatch):
    w2i, i2w = synth.tiny_vocab(97)
    m = ref.MultimodalTransformer(64, 128, 48, 96, 40, w2i, i2w).eval()
    sd = synth.synth_state_dict(m.state_dict(), seed=3)
    m.load_state_dict(sd)
    xi, xli, xa, xla, y_in, _ = synth.synth_multimodal_batch(2, (64, 128), (48, 96), [15, 9], w2i)
    monkeypatch.setattr(m, "apply_teacher_forcing_modality", lambda: modality)
    with torch.no_grad():
        out = m(xi, xli, xa, xla, y_in, apply_teacher_forcing_modality=True)
        mine = restate.multimodal_forward(sd, xi, xli, xa, xla, y_in, modality=modality)
    assert rel(mine, out) < 5e-6


@pytest.mark.parametrize("window", [-1, 5])
def test_greedy_loop(ref, window):
    w2i, i2w = synth.tiny_vocab(97)
    m = ref.Transformer(64, 128, 20, w2i, i2w, attn_window=window).eval()
    sd = synth.synth_state_dict(m.state_dict(), seed=4)
    m.load_state_dict(sd)
    x = torch.rand(1, 1, 64, 128, generator=torch.Generator().manual_seed(2))
    m.Y, m.YHat = [], []
    m.validation_step((x, torch.tensor([[w2i["<sos>"], 5, w2i["<eos>"]]])), 0)
    mem = restate.encode_to_memory(sd, "encoder.", "pos_2d.pe", x)
    toks, vals = restate.greedy_decode(sd, mem, w2i["<sos>"], w2i["<eos>"], 20, attn_window=window)
    assert [w2i[t] for t in m.YHat[0]] == toks
    words, probs = m.get_pred_seq_and_pred_prob_seq(x)
    assert [w2i[t] for t in words] == toks and max(abs(a - b) for a, b in zip(probs, vals)) < 1e-4


def test_state_dict_surface_matches(ref):
    import omr_a2s_multimodal_transformer_b200 as pkg

    w2i, i2w = synth.load_vocab()
    for mixer in ["concat", "attn_img"]:
        a = ref.MultimodalTransformer(361, 4412, 195, 808, 1268, w2i, i2w, mixer_type=mixer).state_dict()
        b = pkg.MultimodalTransformer(361, 4412, 195, 808, 1268, w2i, i2w, mixer_type=mixer).state_dict()
        assert list(a.keys()) == list(b.keys())
        assert all(a[k].shape == b[k].shape and a[k].dtype == b[k].dtype for k in a)
        assert all(torch.equal(a[k], b[k]) for k in a if k.endswith(".pe"))
    a = ref.Transformer(128, 1024, 1268, w2i, i2w).state_dict()
    b = pkg.Transformer(128, 1024, 1268, w2i, i2w).state_dict()
    assert list(a.keys()) == list(b.keys()) and sum(v.numel() for k, v in b.items() if not k.endswith(".pe")) == 10128309


def test_reference_checkpoint_loads_unchanged(ref, tmp_path):
    """a Lightning-style checkpoint of the REAL reference model ({state_dict, hyper_parameters = ctor args}) loads into
    the replacement through load_from_checkpoint, with overrides, exactly as src/test.py:61-62 does"""
    import inspect

    import omr_a2s_multimodal_transformer_b200 as pkg

    w2i, i2w = synth.tiny_vocab(53)
    torch.manual_seed(0)
    for ref_cls, cls, args in [
        (ref.Transformer, pkg.Transformer, dict(max_input_height=64, max_input_width=128, max_seq_len=20, w2i=w2i, i2w=i2w,
                                                attn_window=7, teacher_forcing_prob=0.3)),
        (ref.MultimodalTransformer, pkg.MultimodalTransformer,
         dict(max_img_height=64, max_img_width=128, max_audio_height=48, max_audio_width=96, max_seq_len=20, w2i=w2i, i2w=i2w,
              mixer_type="attn_both", attn_window=-1)),
    ]:
        # same constructor signature, so the reference's hyper_parameters are valid constructor arguments here
        assert list(inspect.signature(ref_cls.__init__).parameters) == list(inspect.signature(cls.__init__).parameters)
        m_ref = ref_cls(**args)
        path = tmp_path / f"{cls.__name__}.ckpt"
        torch.save({"state_dict": m_ref.state_dict(), "hyper_parameters": args}, path)
        m = cls.load_from_checkpoint(str(path), ytest_i2w=i2w)
        sd_ref, sd = m_ref.state_dict(), m.state_dict()
        assert list(sd) == list(sd_ref) and all(torch.equal(sd[k], sd_ref[k]) for k in sd)
        assert m.max_seq_len == 20 and m.attn_window == args["attn_window"] and m.ytest_i2w == i2w
        m.freeze()
        assert not m.training and not any(p.requires_grad for p in m.parameters())


def test_mask_builders_match_reference_over_random_sizes(ref):
    """window / causal / target-padding / memory-padding / concat masks of the restatement against the reference's own
    builders (decoder.py:150-254, model.py:644-675) for random sizes, windows and lengths -- and the product's host-side
    mask builders (kept for API compatibility) against the same"""
    import random

    import omr_a2s_multimodal_transformer_b200 as pkg

    rnd = random.Random(11)
    w2i, i2w = synth.tiny_vocab(31)
    for _ in range(25):
        t, s, b = rnd.randint(1, 40), rnd.randint(1, 30), rnd.randint(1, 4)
        window = rnd.choice([-1, 1, 2, 5, 39, 40, 100])
        rdec = ref.Decoder(31, 64, 31, attn_window=window)
        pdec = pkg.Decoder(31, 64, 31, attn_window=window)
        tgt = torch.randint(0, 31, (b, t), generator=torch.Generator().manual_seed(rnd.randint(0, 10 ** 6)))
        rm, rk = rdec.get_tgt_masks(tgt)
        om, ok = restate.tgt_masks(tgt, window, torch.float32)
        pm, pk = pdec.get_tgt_masks(tgt)
        assert torch.equal(rm, om) and torch.equal(rk, ok) and torch.equal(rm, pm) and torch.equal(rk, pk)
        if window > 0:
            wm = ref.Decoder.create_variable_window_mask(t, window)
            assert torch.equal(wm, restate.window_mask(t, window, torch.float32))
            assert torch.equal(wm, pkg.Decoder.create_variable_window_mask(t, window))
        mem = torch.zeros(b, s, 256)
        lens = torch.tensor([rnd.randint(1, s) for _ in range(b)], dtype=torch.int32)
        for ml in (None, lens, torch.rand(b, s) < 0.3):
            r = rdec.get_memory_key_padding_mask(mem, ml)
            o = restate.memory_key_padding_mask(mem, ml)
            p = pdec.get_memory_key_padding_mask(mem, ml)
            assert (r is None and o is None and p is None) or (r.dtype == o.dtype == p.dtype and torch.equal(r, o) and torch.equal(r, p))
    rmm = ref.MultimodalTransformer(64, 128, 48, 96, 20, w2i, i2w)
    for _ in range(10):
        b, li, la = rnd.randint(1, 4), rnd.randint(1, 20), rnd.randint(1, 20)
        xi, xa = torch.rand(b, li, 256), torch.rand(b, la, 256)
        xli = torch.tensor([rnd.randint(1, li) for _ in range(b)], dtype=torch.int32)
        xla = torch.tensor([rnd.randint(1, la) for _ in range(b)], dtype=torch.int32)
        rx, rl = rmm.mixer_concat(xi, xa, xli, xla)
        ox, ol = restate.mixer_concat(xi, xa, xli, xla)
        assert torch.equal(rx, ox) and rl.dtype == ol.dtype and torch.equal(rl, ol)
