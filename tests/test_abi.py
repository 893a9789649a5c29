"""CPU: the C-ABI shared library builds for sm_100a, loads, and exports every symbol that
include/omr_b200.h declares (no compute calls without a GPU); host-side error behaviour."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "omr_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(omr_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(lib_built):
    from omr_a2s_multimodal_transformer_b200 import _lib

    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 35
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    bound = set(_lib.exported_symbols()) | {"omr_set_tensor_core_path"}
    assert set(syms) <= bound, sorted(set(syms) - bound)
    assert lib.omr_abi_version() == 5


def test_sass_is_sm100a(lib_built):
    import subprocess

    from omr_a2s_multimodal_transformer_b200 import _lib

    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out


def test_no_cpu_fallback_and_clear_errors(lib_built):
    import omr_a2s_multimodal_transformer_b200 as pkg
    from oracle import synth

    w2i, i2w = synth.tiny_vocab(31)
    m = pkg.Transformer(32, 64, 12, w2i, i2w)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 1, 32, 64), None, torch.zeros(1, 3, dtype=torch.long))
    with pytest.raises(ValueError, match="Invalid mixer type"):
        pkg.MultimodalTransformer(32, 64, 32, 64, 12, w2i, i2w, mixer_type="nope")
    with pytest.raises(NotImplementedError):
        pkg.Decoder(31, 12, 31, embedding_dim=256, nhead=8)
    # the steps either side of the model (staging.py) are device-only as well
    from omr_a2s_multimodal_transformer_b200 import staging

    sample = (torch.rand(1, 5, 7), 1, torch.tensor([1, 2, 3]))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        staging.ar_batch_preparation_image([sample], device="cpu")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        staging.compute_ed_metrics([["a"]], [["b"]], device="cpu")
    with pytest.raises(ValueError, match="Vocabularies do not match"):
        pkg.WeightedGreedyDecoder(pkg.Decoder(31, 12, 31), pkg.Decoder(30, 12, 30))


def test_product_never_imports_oracle():
    pkg_dir = os.path.join(ROOT, "omr_a2s_multimodal_transformer_b200")
    for fn in os.listdir(pkg_dir):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg_dir, fn)).read()
            assert "oracle" not in src.replace("# oracle", ""), fn


def test_grad_arena_and_weight_cache_host_logic():
    from omr_a2s_multimodal_transformer_b200.params import GradArena

    ps = [torch.nn.Parameter(torch.zeros(3, 5)), torch.nn.Parameter(torch.zeros(7))]
    arena = GradArena(ps)
    assert arena.attached() and arena.flat.numel() == 128
    ps[0].grad.add_(1.0)
    assert float(arena.flat.sum()) == 15.0
    arena.zero_()
    assert float(ps[0].grad.sum()) == 0.0
    ps[1].grad = None
    assert not arena.attached()
    arena.reattach()
    assert arena.attached()


def test_drop_in_module_aliases_accept_the_callers_keyword_arguments(monkeypatch):
    """alias the reference's module names, then construct the models exactly as
    src/train.py:80-104 does (keyword arguments) -- no edit to the callers"""
    import sys

    import omr_a2s_multimodal_transformer_b200 as b200
    from oracle import synth

    for name, mod in (("src.transformer.model", b200.model), ("src.transformer.encoder", b200.encoder),
                      ("src.transformer.decoder", b200.decoder)):
        monkeypatch.setitem(sys.modules, name, mod)
    for parent in ("src", "src.transformer"):
        if parent not in sys.modules:
            import types

            monkeypatch.setitem(sys.modules, parent, types.ModuleType(parent))
    from src.transformer.model import MultimodalTransformer, Transformer  # noqa: E402  (the callers' import line)

    w2i, i2w = synth.tiny_vocab(31)
    m = MultimodalTransformer(max_img_height=32, max_img_width=64, max_audio_height=32, max_audio_width=64, max_seq_len=12,
                              w2i=w2i, i2w=i2w, mixer_type="concat", attn_window=100, teacher_forcing_prob=0.2,
                              teacher_forcing_modality_prob=0.2)
    u = Transformer(max_input_height=32, max_input_width=64, max_seq_len=12, w2i=w2i, i2w=i2w, attn_window=100,
                    teacher_forcing_prob=0.2)
    assert m.teacher_forcing_modality_prob == 0.2 and u.teacher_forcing_prob == 0.2 and m.decoder.attn_window == 100
    # the sub-module names the late-fusion scripts and the checkpoint splitter reach into
    for attr in ("image_encoder", "audio_encoder", "image_pos_2d", "audio_pos_2d", "decoder"):
        assert hasattr(m, attr)
    for attr in ("encoder", "pos_2d", "decoder", "w2i", "i2w", "ytest_i2w", "padding_idx", "max_seq_len", "Y", "YHat"):
        assert hasattr(u, attr)
    opt = u.configure_optimizers()
    assert isinstance(opt, torch.optim.Optimizer) and opt.param_groups[0]["lr"] == 1e-4
