"""Input staging and validation metrics on the device -- the steps either side of the model.

* ``ar_batch_preparation_unimodal / _image / _audio / _multimodal``: the reference's collate functions
  (``src/data/preprocessing.py:55-144``) with the same names, arguments and return tuples.  The ragged samples of a batch
  are packed into ONE pinned host buffer, copied with one H2D transfer, and padded on the GPU by ``omr_pad_collate`` /
  ``omr_pad_transcripts`` (image background 1.0, spectrogram background 0.0); the returned tensors are on the device.
* ``number_of_frames``: ``get_number_of_frames`` (``src/data/ar_dataset.py:439-442``), computed by the same kernel.
* ``compute_ed_metrics``: Sym-ER / Seq-ER (``src/utils/metrics.py:52-88``) with the Levenshtein distances of all pairs
  computed by one launch of ``omr_levenshtein``.

There is no CPU fallback: without a CUDA device these functions raise."""
from __future__ import annotations

from typing import Dict, Hashable, List, Optional, Sequence, Tuple

import torch

from . import ops

HEIGHT_REDUCTION, WIDTH_REDUCTION = 16, 8  # src/data/ar_dataset.py:22-23 ... the encoder's total strides


def _device(device) -> torch.device:
    dev = torch.device("cuda:0" if device is None else device)
    if dev.type != "cuda":
        raise RuntimeError(f"staging runs only on CUDA (got {dev}); there is no CPU fallback")
    return dev


def _pack(samples: Sequence[torch.Tensor], dtype: torch.dtype, dev: torch.device):
    """ragged tensors -> (flat device buffer, int64 device offsets [n+1]) through one pinned staging buffer"""
    sizes = [int(s.numel()) for s in samples]
    offs = [0]
    for n in sizes:
        offs.append(offs[-1] + n)
    on_dev = all(s.is_cuda for s in samples)
    if on_dev:
        flat = torch.cat([s.reshape(-1).to(dtype) for s in samples]) if samples else torch.empty(0, dtype=dtype, device=dev)
    else:
        host = torch.empty((offs[-1],), dtype=dtype, pin_memory=True)
        for s, o, n in zip(samples, offs, sizes):
            host[o:o + n].copy_(s.reshape(-1))
        flat = host.to(dev, non_blocking=True)
    offsets = torch.tensor(offs, dtype=torch.int64).to(dev, non_blocking=True)
    return flat, offsets


def pad_batch_inputs(x: Sequence[torch.Tensor], pad_value: float = 0.0, dtype: torch.dtype = torch.float32,
                     device=None) -> torch.Tensor:
    """list of [1,h,w] samples -> [B,1,max_h,max_w] on the device (preprocessing.py:55-74)"""
    if dtype != torch.float32:
        raise NotImplementedError("pad_batch_inputs: the reference only ever asks for float32")
    return _pad_inputs(x, pad_value, _device(device))[0]


def _pad_inputs(x, pad_value, dev):
    for s in x:
        if s.dim() != 3 or s.shape[0] != 1:
            raise ValueError(f"expected [1,h,w] samples, got {tuple(s.shape)}")
    hs = [int(s.shape[1]) for s in x]
    ws = [int(s.shape[2]) for s in x]
    flat, offsets = _pack(x, torch.float32, dev)
    hw = torch.tensor([hs, ws], dtype=torch.int32).to(dev, non_blocking=True)
    return ops.pad_collate(flat, offsets, hw[0], hw[1], max(hs), max(ws), pad_value, HEIGHT_REDUCTION, WIDTH_REDUCTION)


def number_of_frames(x: Sequence[torch.Tensor], device=None) -> torch.Tensor:
    """int32 [B]: ceil(h/16) * ceil(w/8) of every sample (ar_dataset.py:439-442)"""
    return _pad_inputs(x, 0.0, _device(device))[1]


def pad_batch_transcripts_shifted(y: Sequence[torch.Tensor], device=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """transcripts (with <sos> ... <eos>) -> (y_in, y_out) = padded transcript[:-1] / transcript[1:], int64 [B,maxlen-1]"""
    dev = _device(device)
    flat, offsets = _pack(y, torch.int64, dev)
    t = max(int(s.numel()) for s in y) - 1
    return ops.pad_transcripts(flat, offsets, max(t, 0), 0)


def ar_batch_preparation_unimodal(batch, pad_value: float = 0.0, device=None):
    x, xl, y = zip(*batch)
    dev = _device(device)
    xp, _ = _pad_inputs(x, pad_value, dev)
    xl = torch.tensor(xl, dtype=torch.int32).to(dev, non_blocking=True)
    y_in, y_out = pad_batch_transcripts_shifted(y, dev)
    return xp, xl, y_in, y_out


def ar_batch_preparation_image(batch, device=None):
    return ar_batch_preparation_unimodal(batch, pad_value=1.0, device=device)  # white score background


def ar_batch_preparation_audio(batch, device=None):
    return ar_batch_preparation_unimodal(batch, device=device)  # black spectrogram background


def ar_batch_preparation_multimodal(batch, device=None):
    xi, xli, xa, xla, y = zip(*batch)
    dev = _device(device)
    xip, _ = _pad_inputs(xi, 1.0, dev)
    xap, _ = _pad_inputs(xa, 0.0, dev)
    xli = torch.tensor(xli, dtype=torch.int32).to(dev, non_blocking=True)
    xla = torch.tensor(xla, dtype=torch.int32).to(dev, non_blocking=True)
    y_in, y_out = pad_batch_transcripts_shifted(y, dev)
    return xip, xli, xap, xla, y_in, y_out


def edit_distances(y_true: Sequence[Sequence[Hashable]], y_pred: Sequence[Sequence[Hashable]], device=None):
    """-> (int32 [P] distances, int64 [3] = {sum of distances, sum of truth lengths, #pairs with an error}) on the device"""
    if len(y_true) != len(y_pred):
        raise ValueError("y_true and y_pred differ in length")
    dev = _device(device)
    ids: Dict[Hashable, int] = {}

    def enc(seqs):
        offs, flat = [0], []
        for s in seqs:
            flat.extend(ids.setdefault(w, len(ids)) for w in s)
            offs.append(len(flat))
        return (torch.tensor(flat, dtype=torch.int64).to(dev, non_blocking=True),
                torch.tensor(offs, dtype=torch.int64).to(dev, non_blocking=True))

    t, to = enc(y_true)
    h, ho = enc(y_pred)
    max_len = max([len(s) for s in y_true] + [0])
    return ops.levenshtein(t, to, h, ho, max_len)


def compute_ed_metrics(y_true: List[List[str]], y_pred: List[List[str]], device=None) -> Dict[str, float]:
    _, sums = edit_distances(y_true, y_pred, device)
    ed, length, wrong = (int(v) for v in sums.tolist())
    return {"sym-er": 100.0 * ed / length, "seq-er": 100.0 * wrong / len(y_pred)}
