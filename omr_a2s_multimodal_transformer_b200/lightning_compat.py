"""``LightningModule`` base for the model classes.

The reference models derive from ``lightning.pytorch.LightningModule`` (reference model.py:7,54,358)
and are driven by ``lightning.Trainer`` from ``src/train.py`` / ``src/test.py``.  When Lightning is
installed the real base class is used, so the replacement drops into those callers unchanged.  When
it is not (this build image), a minimal stand-in supplies the few members the model surface needs
(``save_hyperparameters``, ``log``, ``device``, ``freeze``, ``load_from_checkpoint``).
"""
from __future__ import annotations

import inspect
from typing import Any, Dict

import torch
import torch.nn as nn

try:  # pragma: no cover - depends on the environment
    from lightning.pytorch import LightningModule  # type: ignore

    # a partial stand-in registered under that name (e.g. a test stub) does not count as Lightning
    HAVE_LIGHTNING = all(hasattr(LightningModule, a) for a in ("load_from_checkpoint", "save_hyperparameters", "freeze"))
except Exception:  # pragma: no cover
    HAVE_LIGHTNING = False

if not HAVE_LIGHTNING:

    class LightningModule(nn.Module):  # type: ignore
        def __init__(self) -> None:
            super().__init__()
            self.hparams: Dict[str, Any] = {}
            self.logged: Dict[str, Any] = {}

        def save_hyperparameters(self, *args, **kwargs) -> None:
            frame = inspect.currentframe().f_back
            init_args = {k: v for k, v in frame.f_locals.items() if k not in ("self", "__class__")}
            self.hparams = dict(init_args)

        def log(self, name: str, value: Any, *args, **kwargs) -> None:
            self.logged[name] = value

        @property
        def device(self) -> torch.device:
            try:
                return next(self.parameters()).device
            except StopIteration:
                return torch.device("cpu")

        def freeze(self) -> None:
            for p in self.parameters():
                p.requires_grad_(False)
            self.eval()

        def unfreeze(self) -> None:
            for p in self.parameters():
                p.requires_grad_(True)
            self.train()

        @classmethod
        def load_from_checkpoint(cls, checkpoint_path, map_location=None, strict: bool = True, **overrides):
            ckpt = torch.load(checkpoint_path, map_location=map_location or "cpu", weights_only=False)
            hparams = dict(ckpt.get("hyper_parameters", {}))
            hparams.update(overrides)
            model = cls(**hparams)
            model.load_state_dict(ckpt["state_dict"], strict=strict)
            return model

        def on_save_checkpoint_dict(self) -> Dict[str, Any]:
            """What ``Trainer.save_checkpoint`` would write for this module (state + hyper-parameters)."""
            return {"state_dict": self.state_dict(), "hyper_parameters": dict(self.hparams)}
