"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the steps either side of the model.

Plain torch / Python, each function citing the reference lines it follows; pinned against the reference's own code in
``tests/test_oracle_pin.py`` (collate and metrics: the imported ``src.data.preprocessing`` / ``src.utils.metrics``;
late fusion: the reference's ``weighted_prediction`` function body executed from its source file).  Nothing in the
product path imports this module."""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn.functional as F

from . import restate

HEIGHT_REDUCTION, WIDTH_REDUCTION = 16, 8  # src/data/ar_dataset.py:22-23


def pad_batch_inputs(x: Sequence[torch.Tensor], pad_value: float = 0.0) -> torch.Tensor:
    """src/data/preprocessing.py:55-74: right/bottom padding to the batch maximum, float32"""
    mw = max(int(s.shape[2]) for s in x)
    mh = max(int(s.shape[1]) for s in x)
    out = torch.full((len(x), 1, mh, mw), float(pad_value), dtype=torch.float32)
    for b, s in enumerate(x):
        out[b, :, : s.shape[1], : s.shape[2]] = s.to(torch.float32)
    return out


def number_of_frames(x: torch.Tensor) -> int:
    """src/data/ar_dataset.py:439-442"""
    return math.ceil(x.shape[1] / HEIGHT_REDUCTION) * math.ceil(x.shape[2] / WIDTH_REDUCTION)


def pad_batch_transcripts(y: Sequence[torch.Tensor]) -> torch.Tensor:
    """src/data/preprocessing.py:77-81 (zero padding on the right, int64)"""
    m = max(int(s.shape[0]) for s in y)
    return torch.stack([F.pad(s, (0, m - s.shape[0])) for s in y]).to(torch.int64)


def ar_batch_preparation_unimodal(batch, pad_value: float = 0.0):
    """src/data/preprocessing.py:84-100"""
    x, xl, y = zip(*batch)
    return (pad_batch_inputs(x, pad_value), torch.tensor(xl, dtype=torch.int32), pad_batch_transcripts([t[:-1] for t in y]),
            pad_batch_transcripts([t[1:] for t in y]))


def ar_batch_preparation_multimodal(batch):
    """src/data/preprocessing.py:118-144: image background 1.0, spectrogram background 0.0"""
    xi, xli, xa, xla, y = zip(*batch)
    return (pad_batch_inputs(xi, 1.0), torch.tensor(xli, dtype=torch.int32), pad_batch_inputs(xa, 0.0),
            torch.tensor(xla, dtype=torch.int32), pad_batch_transcripts([t[:-1] for t in y]),
            pad_batch_transcripts([t[1:] for t in y]))


def levenshtein(a: Sequence, b: Sequence) -> int:
    """src/utils/metrics.py:56-73 (two-row dynamic programme, unit costs)"""
    n, m = len(a), len(b)
    if n > m:
        a, b, n, m = b, a, m, n
    cur = list(range(n + 1))
    for i in range(1, m + 1):
        prev, cur = cur, [i] + [0] * n
        for j in range(1, n + 1):
            cur[j] = min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (a[j - 1] != b[i - 1]))
    return cur[n]


def compute_ed_metrics(y_true: List[List], y_pred: List[List]) -> Dict[str, float]:
    """src/utils/metrics.py:52-88"""
    ed_acc = len_acc = lab = 0
    for t, h in zip(y_true, y_pred):
        ed = levenshtein(t, h)
        ed_acc += ed
        len_acc += len(t)
        lab += ed > 0
    return {"sym-er": 100.0 * ed_acc / len_acc, "seq-er": 100.0 * lab / len(y_pred)}


@torch.no_grad()
def weighted_greedy_decode(sd_img, sd_aud, mem_img: torch.Tensor, mem_aud: torch.Tensor, sos: int, eos: int, max_seq_len: int,
                           alpha: float = 0.5, nhead: int = 4, num_layers: int = 8) -> Tuple[List[int], List[float]]:
    """src/multimodal/weighted_multimodal/test.py:44-70: both decoders re-run on the growing shared prefix, softmax of
    each last-position logit vector, alpha mix, argmax; EOS is emitted and ends the loop.  -> (tokens, mixed probs)"""
    assert mem_img.shape[0] == 1 and mem_aud.shape[0] == 1, "Inference only supports batch_size = 1"
    y_in = torch.tensor([[sos]], dtype=torch.int64)
    toks: List[int] = []
    vals: List[float] = []
    for _ in range(max_seq_len):
        pi = restate.decoder_forward(sd_img, "decoder.", y_in, mem_img, None, -1, nhead, num_layers)[0, :, -1].softmax(dim=-1)
        pa = restate.decoder_forward(sd_aud, "decoder.", y_in, mem_aud, None, -1, nhead, num_layers)[0, :, -1].softmax(dim=-1)
        p = alpha * pi + (1 - alpha) * pa
        tok = int(p.argmax(dim=-1))
        toks.append(tok)
        vals.append(float(p[tok]))
        if tok == eos:
            break
        y_in = torch.cat([y_in, torch.tensor([[tok]], dtype=torch.int64)], dim=1)
    return toks, vals
