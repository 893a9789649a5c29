"""TEST INFRASTRUCTURE ONLY.  Loader for the REAL reference modules (container only).

The reference (``/root/reference``, read-only) is pure Python on top of torch, but
``src/transformer/model.py`` imports packages that are absent from this image (``lightning``,
``torchinfo``, ``librosa``, ``music21``, ``midi2audio``, ``pyMV2H``).  This module registers
minimal stand-ins for those packages in ``sys.modules`` and then imports the reference's own,
unmodified ``Encoder`` / ``Decoder`` / ``Transformer`` / ``MultimodalTransformer``.

It exists to (a) pin ``oracle/restate.py`` against the real code and (b) generate the golden
vectors under ``tests/golden``.  ``/root/reference`` does not exist on the GPU box, so nothing
that runs there may call :func:`load_reference`; use :func:`reference_available` to skip.
"""
from __future__ import annotations

import os
import sys
import types
from types import SimpleNamespace

import torch
import torch.nn as nn

def _find_reference_root() -> str:
    """the read-only reference checkout in the build container, else the git-ignored verbatim copy that
    oracle/build_ref.py makes under oracle/_ref (the only form in which the reference reaches the GPU box)"""
    cands = [os.environ.get("OMR_REFERENCE_ROOT"), "/root/reference",
             os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")]
    for c in cands:
        if c and os.path.isfile(os.path.join(c, "src", "transformer", "model.py")):
            return c
    return cands[1]


REFERENCE_ROOT = _find_reference_root()


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "transformer", "model.py"))


class _LightningModuleStub(nn.Module):
    """nn.Module carrying just the LightningModule API the reference model touches."""

    def save_hyperparameters(self, *a, **k):
        return None

    def log(self, *a, **k):
        return None

    @property
    def device(self):
        try:
            return next(self.parameters()).device
        except StopIteration:
            return torch.device("cpu")

    def freeze(self):
        for p in self.parameters():
            p.requires_grad_(False)
        self.eval()


class _LightningDataModuleStub:
    def __init__(self, *a, **k):
        pass


def _module(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def _install_stubs() -> None:
    try:  # prefer the real thing when it exists
        import lightning.pytorch  # noqa: F401
    except Exception:
        lp = _module(
            "lightning.pytorch",
            LightningModule=_LightningModuleStub,
            LightningDataModule=_LightningDataModuleStub,
        )
        _module("lightning", pytorch=lp)
    try:
        import torchinfo  # noqa: F401
    except Exception:
        _module("torchinfo", summary=lambda *a, **k: None)
    for name, attrs in [
        ("librosa", {}),
        ("midi2audio", {"FluidSynth": object}),
        ("music21", {"converter": SimpleNamespace()}),
        ("pyMV2H", {}),
        ("pyMV2H.converter", {}),
        ("pyMV2H.converter.midi_converter", {"MidiConverter": object}),
        ("pyMV2H.metrics", {}),
        ("pyMV2H.metrics.mv2h", {"mv2h": lambda *a, **k: None}),
        ("pyMV2H.utils", {}),
        ("pyMV2H.utils.music", {"Music": object}),
        ("pyMV2H.utils.mv2h", {"MV2H": object}),
    ]:
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                _module(name, **attrs)


_CACHE = None


def load_reference() -> SimpleNamespace:
    """Import the unmodified reference modules; returns a namespace with the classes."""
    global _CACHE
    if _CACHE is not None:
        return _CACHE
    if not reference_available():
        raise RuntimeError(
            f"reference not found at {REFERENCE_ROOT}; the oracle shim only works in the build container"
        )
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # src.data.preprocessing creates ./joblib_cache in the cwd at import: do it from a scratch dir
    cwd = os.getcwd()
    scratch = os.environ.get("OMR_ORACLE_SCRATCH", "/tmp/omr_oracle_scratch")
    os.makedirs(scratch, exist_ok=True)
    os.chdir(scratch)
    try:
        # our repo may shadow the name "src" if a drop-in shim package was installed
        for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
            del sys.modules[k]
        import src.transformer.encoder as enc
        import src.transformer.decoder as dec
        import src.transformer.model as mod
    finally:
        os.chdir(cwd)
    _CACHE = SimpleNamespace(
        encoder=enc,
        decoder=dec,
        model=mod,
        Encoder=enc.Encoder,
        ConvBlock=enc.ConvBlock,
        DSCBlock=enc.DSCBlock,
        Decoder=dec.Decoder,
        PositionalEncoding1D=dec.PositionalEncoding1D,
        PositionalEncoding2D=mod.PositionalEncoding2D,
        CrossAttention=mod.CrossAttention,
        Transformer=mod.Transformer,
        MultimodalTransformer=mod.MultimodalTransformer,
    )
    return _CACHE
