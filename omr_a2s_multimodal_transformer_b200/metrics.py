"""Symbol / sequence error rates used by ``on_validation_epoch_end`` (reference
``src/utils/metrics.py:52-88``): token-level Levenshtein distance summed over the set divided by the
total reference length (Sym-ER, %), and the share of sequences with at least one error (Seq-ER, %).
Host-side bookkeeping on short token lists; MV2H (music21 / pyMV2H) is outside the hot path."""
from __future__ import annotations

from typing import Dict, List, Sequence


def edit_distance(a: Sequence, b: Sequence) -> int:
    if len(a) < len(b):
        a, b = b, a
    prev = list(range(len(b) + 1))
    for i, x in enumerate(a, 1):
        cur = [i]
        for j, y in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (x != y)))
        prev = cur
    return prev[-1]


def compute_metrics(y_true: List[List[str]], y_pred: List[List[str]], compute_mv2h: bool = False) -> Dict[str, float]:
    if compute_mv2h:
        raise NotImplementedError("MV2H needs music21/pyMV2H and is not part of the accelerated path")
    ed_total = length_total = wrong = 0
    for t, h in zip(y_true, y_pred):
        ed = edit_distance(t, h)
        ed_total += ed
        length_total += len(t)
        wrong += ed > 0
    return {"sym-er": 100.0 * ed_total / length_total, "seq-er": 100.0 * wrong / len(y_pred)}
