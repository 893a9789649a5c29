"""TEST INFRASTRUCTURE ONLY.  Deterministic synthetic weights and GrandStaff-shaped batches.

Weights are generated PER STATE-DICT KEY from a hash of the key, so the same values can be loaded
into the real reference modules (build container), into ``oracle/restate.py`` and into the CUDA
modules (GPU box) without depending on module construction order or on torch's default init RNG
stream.  Batches follow SURVEY.md section 8(d) and the collate contract of the reference
(``src/data/preprocessing.py:85-144``: image pad value 1.0, audio pad value 0.0, lengths int32,
``y_in``/``y_out`` int64 padded with ``<PAD>``=0).
"""
from __future__ import annotations

import json
import math
import os
import zlib
from typing import Dict, List, Optional, Tuple

import torch

REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VOCAB_PATH = os.path.join(REPO_ROOT, "grandstaff", "vocabs", "ar_w2i_kern.json")
MAXLEN_PATH = os.path.join(REPO_ROOT, "grandstaff", "max_lens", "ImgDist_ar_w2i_kern.json")

HEIGHT_REDUCTION = 16  # reference src/transformer/encoder.py:8
WIDTH_REDUCTION = 8  # reference src/transformer/encoder.py:9


def load_vocab() -> Tuple[Dict[str, int], Dict[int, str]]:
    with open(VOCAB_PATH) as f:
        w2i = json.load(f)
    i2w = {v: k for k, v in w2i.items()}
    return w2i, i2w


def load_max_lens() -> Dict[str, int]:
    with open(MAXLEN_PATH) as f:
        return json.load(f)


def tiny_vocab(n: int = 97) -> Tuple[Dict[str, int], Dict[int, str]]:
    """Small vocabulary with the same special-token conventions (<PAD>=0, <eos>, <sos> near the end)."""
    w2i = {"<PAD>": 0}
    for i in range(1, n - 2):
        w2i[f"tok{i}"] = i
    w2i["<eos>"] = n - 2
    w2i["<sos>"] = n - 1
    return w2i, {v: k for k, v in w2i.items()}


def _gen_for(key: str, seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def synth_state_dict(template: Dict[str, torch.Tensor], seed: int = 0, padding_idx: int = 0) -> Dict[str, torch.Tensor]:
    """Fill every tensor of ``template`` (a ``state_dict()``) with key-hashed deterministic values.

    ``*.pe`` buffers are left untouched (they are closed-form constants).
    """
    out: Dict[str, torch.Tensor] = {}
    for key in template:
        t = template[key]
        if key.endswith(".pe") or not torch.is_floating_point(t):
            out[key] = t.clone()
            continue
        g = _gen_for(key, seed)
        shape = tuple(t.shape)
        leaf = key.split(".")[-2] if "." in key else ""
        if leaf.startswith("norm") and key.endswith("weight") and t.dim() == 1:
            v = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif leaf.startswith("norm") and key.endswith("bias"):
            v = 0.1 * torch.randn(shape, generator=g)
        elif key.endswith("embedding.weight"):
            v = torch.randn(shape, generator=g)
            v[padding_idx].zero_()
        elif t.dim() >= 2:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            b = 1.0 / math.sqrt(max(fan_in, 1))
            v = (torch.rand(shape, generator=g) * 2 - 1) * b
        else:
            v = (torch.rand(shape, generator=g) * 2 - 1) * 0.05
        out[key] = v.to(t.dtype)
    return out


def n_frames(h: int, w: int) -> int:
    """reference src/data/ar_dataset.py:439-442"""
    return math.ceil(h / HEIGHT_REDUCTION) * math.ceil(w / WIDTH_REDUCTION)


def synth_tokens(
    batch: int,
    total_lens: List[int],
    w2i: Dict[str, int],
    g: torch.Generator,
) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """``y = [<sos>] + uniform non-special ids + [<eos>]``, 0-padded; returns (y, y_in, y_out)."""
    V = len(w2i)
    sos, eos = w2i["<sos>"], w2i["<eos>"]
    tmax = max(total_lens)
    y = torch.zeros(batch, tmax, dtype=torch.int64)
    specials = {0, sos, eos}
    allowed = torch.tensor([i for i in range(V) if i not in specials], dtype=torch.int64)
    for b, n in enumerate(total_lens):
        assert n >= 2
        idx = torch.randint(0, allowed.numel(), (n - 2,), generator=g)
        y[b, 0] = sos
        y[b, 1 : n - 1] = allowed[idx]
        y[b, n - 1] = eos
    return y, y[:, :-1].contiguous(), y[:, 1:].contiguous()


def synth_unimodal_batch(
    batch: int,
    height: int,
    width: int,
    total_lens: List[int],
    w2i: Dict[str, int],
    seed: int = 1,
    pad_value: float = 1.0,
    frame_lens: Optional[List[int]] = None,
):
    """(x, xl, y_in, y_out) as produced by ``ar_batch_preparation_unimodal``."""
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(batch, 1, height, width, generator=g)
    full = n_frames(height, width)
    if frame_lens is None:
        frame_lens = [max(1, full - (i * full) // (2 * batch)) for i in range(batch)]
    xl = torch.tensor(frame_lens, dtype=torch.int32)
    _, y_in, y_out = synth_tokens(batch, total_lens, w2i, g)
    return x, xl, y_in, y_out


def synth_multimodal_batch(
    batch: int,
    img_hw: Tuple[int, int],
    aud_hw: Tuple[int, int],
    total_lens: List[int],
    w2i: Dict[str, int],
    seed: int = 1,
    img_frame_lens: Optional[List[int]] = None,
    aud_frame_lens: Optional[List[int]] = None,
):
    """(xi, xli, xa, xla, y_in, y_out) as produced by ``ar_batch_preparation_multimodal``."""
    g = torch.Generator().manual_seed(seed)
    xi = torch.rand(batch, 1, *img_hw, generator=g)
    xa = torch.rand(batch, 1, *aud_hw, generator=g)
    li, la = n_frames(*img_hw), n_frames(*aud_hw)
    if img_frame_lens is None:
        img_frame_lens = [max(1, li - (i * li) // (2 * batch)) for i in range(batch)]
    if aud_frame_lens is None:
        aud_frame_lens = [max(1, la - ((batch - 1 - i) * la) // (2 * batch)) for i in range(batch)]
    xli = torch.tensor(img_frame_lens, dtype=torch.int32)
    xla = torch.tensor(aud_frame_lens, dtype=torch.int32)
    _, y_in, y_out = synth_tokens(batch, total_lens, w2i, g)
    return xi, xli, xa, xla, y_in, y_out


def ragged_lens(batch: int, lo: int, hi: int, seed: int = 7) -> List[int]:
    g = torch.Generator().manual_seed(seed)
    v = torch.randint(lo, hi + 1, (batch,), generator=g).tolist()
    v[0] = hi  # at least one full-length row so T == hi - 1
    return v
