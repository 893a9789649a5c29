"""Token-level late fusion of an image model and an audio model (reference
``src/multimodal/weighted_multimodal/test.py:21-70``, ``weighted_prediction``), on the batched KV-cached decoders.

``weighted_prediction`` keeps the reference's signature and batch-1 contract (it returns the word list of the one
sample); ``weighted_prediction_batch`` decodes a whole batch in lock-step and returns the device tensors."""
from __future__ import annotations

from typing import List, Tuple

import torch

from .greedy import BatchedGreedyDecoder, WeightedGreedyDecoder
from .model import EOS_TOKEN, SOS_TOKEN, Transformer


@torch.no_grad()
def weighted_prediction_batch(xi: torch.Tensor, xa: torch.Tensor, img_model: Transformer, audio_model: Transformer,
                              alpha: float = 0.5, max_steps=None, stop_at_eos: bool = True,
                              use_graph: bool = True) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """xi [B,1,H,W], xa [B,1,F,T] -> (tokens [B,steps], mixed probabilities [B,steps], lengths [B]) on the device"""
    if img_model.w2i != audio_model.w2i:
        raise AssertionError("Vocabularies do not match")  # test.py:140
    mem_i = img_model.encode(xi.to(img_model.device))  # encoder + 2-D PE + flatten/permute (test.py:28-36)
    mem_a = audio_model.encode(xa.to(audio_model.device))
    runner = WeightedGreedyDecoder(img_model.decoder, audio_model.decoder, img_model.compute_dtype)
    w2i = img_model.w2i
    return runner.decode(mem_i, mem_a, w2i[SOS_TOKEN], w2i[EOS_TOKEN], img_model.padding_idx, alpha=alpha,
                         max_steps=max_steps, stop_at_eos=stop_at_eos, use_graph=use_graph)


@torch.no_grad()
def weighted_prediction(xi: torch.Tensor, xa: torch.Tensor, img_model: Transformer, audio_model: Transformer,
                        alpha: float = 0.5) -> List[str]:
    assert xi.size(0) == 1, "Inference only supports batch_size = 1"  # test.py:39
    toks, vals, lens = weighted_prediction_batch(xi, xa, img_model, audio_model, alpha)
    seqs, _ = BatchedGreedyDecoder.to_lists(toks, vals, lens)
    return [img_model._i2w(t) for t in seqs[0]]  # both models share the vocabulary (test.py:63)
