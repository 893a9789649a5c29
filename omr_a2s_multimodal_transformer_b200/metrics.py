"""Symbol / sequence error rates (reference ``src/utils/metrics.py``): ``compute_metrics(y_true, y_pred)`` with the
reference's signature, served by the Levenshtein kernel (``staging.compute_ed_metrics`` -> ``omr_levenshtein``, one CTA
per (truth, prediction) pair).  Like every other entry point of this package it needs a CUDA device; there is no CPU
implementation here (the tests keep their own CPU restatement).  MV2H (music21 /
pyMV2H, ``src/utils/metrics.py:91-175``) is outside the accelerated path."""
from __future__ import annotations

from typing import Dict, List


def compute_metrics(y_true: List[List[str]], y_pred: List[List[str]], compute_mv2h: bool = False, device=None) -> Dict[str, float]:
    if compute_mv2h:
        raise NotImplementedError("MV2H needs music21/pyMV2H and is not part of the accelerated path")
    from .staging import compute_ed_metrics

    return compute_ed_metrics(y_true, y_pred, device=device)
