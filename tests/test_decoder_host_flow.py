"""CPU: the decoder's HOST logic (which C-ABI operators it issues, with which shapes, in forward and in the hand-composed
backward) executed without a GPU: every ``ops`` function the decoder stack calls is replaced by a shape-checking stub.
Covers eval / train mode and the opt-in fused links (OMR_FUSE_DECODER_LINKS=1: dropout folded into the residual
LayerNorms, ReLU + dropout backward of the FFN in one kernel).  Numerics are the business of the -m gpu tests."""
from collections import Counter

import pytest
import torch


@pytest.fixture
def stub_ops(monkeypatch):
    from omr_a2s_multimodal_transformer_b200 import ops

    calls = []

    def embed_pe_fwd(tgt, table, pe, pos0=0, pos_dev=None, out=None):
        calls.append("embed_pe_fwd")
        return torch.zeros(tgt.shape[0], tgt.shape[1], table.shape[1])

    def dropout(x, p, seed, channelwise=False, inplace=False):
        calls.append("dropout")
        return x if inplace else x.clone()

    def linear_fwd(x2d, w, bias=None, relu=False, out=None):
        calls.append("linear_fwd")
        assert x2d.dim() == 2 and x2d.shape[1] == w.shape[1]
        return torch.zeros(x2d.shape[0], w.shape[0]) if out is None else out

    def attn_fwd(q, qo, k, ko, v, vo, spec):
        calls.append("attn_fwd")
        return torch.zeros(q.shape[0], q.shape[1], spec.H * spec.hd), torch.zeros(q.shape[0], spec.H, q.shape[1])

    def _ln(x, res, save):
        assert x.shape == res.shape
        return x.clone(), (x.clone() if save else None), (torch.zeros(x.numel() // x.shape[-1], 2) if save else None)

    def add_layernorm_fwd(x, res, g, bta, eps, save):
        calls.append("add_layernorm_fwd")
        return _ln(x, res, save)

    def dropout_add_layernorm_fwd(x, res, g, bta, eps, save, p, seed):
        calls.append("dropout_add_layernorm_fwd")
        return _ln(x, res, save)

    def layernorm_bwd(dy, s, stats, gamma, dgamma, dbeta):
        calls.append("layernorm_bwd")
        assert dy.shape == s.shape and dgamma.shape == gamma.shape
        return dy.clone()

    def layernorm_bwd_dropout(dy, s, stats, gamma, dgamma, dbeta, p, seed, dbias=None):
        calls.append("layernorm_bwd_dropout")
        assert dy.shape == s.shape and (dbias is None or dbias.shape == gamma.shape)
        return dy.clone(), dy.clone()

    def linear_dgrad(dy2d, w, out=None):
        calls.append("linear_dgrad")
        assert dy2d.shape[1] == w.shape[0]
        return torch.zeros(dy2d.shape[0], w.shape[1])

    def linear_wgrad(x2d, dy2d, dw, db=None, accumulate=True):
        calls.append("linear_wgrad")
        assert x2d.shape[0] == dy2d.shape[0] and tuple(dw.shape) == (dy2d.shape[1], x2d.shape[1])
        assert db is None or tuple(db.shape) == (dy2d.shape[1],)

    def relu_bwd(y, dy, inplace=False):
        calls.append("relu_bwd")
        assert y.shape == dy.shape
        return dy

    def mask_scale(dx, mask, scale):
        calls.append("mask_scale")
        assert dx.shape == mask.shape and abs(scale - 1 / 0.9) < 1e-6
        return dx

    stubs = dict(embed_pe_fwd=embed_pe_fwd, dropout=dropout, linear_fwd=linear_fwd, attn_fwd=attn_fwd,
                 add_layernorm_fwd=add_layernorm_fwd, dropout_add_layernorm_fwd=dropout_add_layernorm_fwd,
                 layernorm_bwd=layernorm_bwd, layernorm_bwd_dropout=layernorm_bwd_dropout, linear_dgrad=linear_dgrad,
                 linear_wgrad=linear_wgrad, relu_bwd=relu_bwd, mask_scale=mask_scale,
                 gemm=lambda *a, **k: calls.append("gemm"), attn_bwd=lambda *a, **k: calls.append("attn_bwd"),
                 embed_bwd=lambda *a, **k: calls.append("embed_bwd"))
    for name, fn in stubs.items():
        assert hasattr(ops, name), name
        monkeypatch.setattr(ops, name, fn)
    return calls


@pytest.mark.parametrize("training", [False, True])
@pytest.mark.parametrize("fuse", ["0", "1"])
def test_decoder_stack_issues_the_expected_operators(stub_ops, monkeypatch, training, fuse):
    import omr_a2s_multimodal_transformer_b200 as pkg

    monkeypatch.setenv("OMR_FUSE_DECODER_LINKS", fuse)
    layers = 2
    dec = pkg.Decoder(31, 12, 31, num_transformer_layers=layers)
    b, t, s, d = 2, 5, 7, 256
    tgt, mem = torch.randint(1, 31, (b, t)), torch.zeros(b, s, d)
    tape = []
    y = dec._run_stack(tgt, mem, None, None, torch.float32, tape, training)
    assert y.shape == (b, t, d)
    fwd = Counter(stub_ops)
    st = {"g": torch.zeros(b, t, d), "dmem": None, "need_dmem": True, "side": None, "keep": []}
    while tape:
        tape.pop()(st)
    assert st["dmem"].shape == (b, s, d) and st["g"].shape == (b, t, d)
    c = Counter(stub_ops)
    # forward: 7 projections, 2 attentions, 3 residual LayerNorms per layer
    assert fwd["linear_fwd"] == 7 * layers and fwd["attn_fwd"] == 2 * layers
    assert fwd["add_layernorm_fwd"] + fwd["dropout_add_layernorm_fwd"] == 3 * layers
    # backward: one weight gradient per projection, the chain's data gradients, both attention backwards
    assert c["linear_wgrad"] == 7 * layers and c["attn_bwd"] == 2 * layers and c["embed_bwd"] == 1
    assert c["linear_dgrad"] == 3 * layers and c["gemm"] == 4 * layers
    if not training:
        assert c["dropout"] == 0 and c["dropout_add_layernorm_fwd"] == 0 and c["mask_scale"] == 0
    elif fuse == "0":
        # embedding (fwd + bwd) + per layer: 4 forward dropouts, 4 backward ones
        assert c["dropout"] == 2 + 8 * layers and c["relu_bwd"] == layers and c["mask_scale"] == 0
    else:
        # only the embedding dropout and the FFN's forward dropout remain separate kernels: 7 links fewer per layer
        assert c["dropout"] == 2 + layers and c["dropout_add_layernorm_fwd"] == 3 * layers
        assert c["layernorm_bwd_dropout"] == 3 * layers and c["mask_scale"] == layers and c["relu_bwd"] == 0


def test_decoder_stack_without_a_tape_runs_the_inference_path(stub_ops):
    import omr_a2s_multimodal_transformer_b200 as pkg

    dec = pkg.Decoder(31, 12, 31, num_transformer_layers=1)
    y = dec._run_stack(torch.randint(1, 31, (1, 3)), torch.zeros(1, 4, 256), None, None, torch.float32, None, False)
    assert y.shape == (1, 3, 256) and "layernorm_bwd" not in stub_ops
