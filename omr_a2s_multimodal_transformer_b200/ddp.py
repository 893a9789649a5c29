"""Data-parallel training over one process per GPU (``torch.distributed``, NCCL over NVLink 5 / NVSwitch).

The reference has no distributed code at all; what a multi-GPU run of its
``src/train.py`` would get from Lightning is stock DDP: replicated parameters, batch sharding, gradient
mean over ranks.  This module provides exactly that for the CUDA path, built around the way its
backward works:

* all gradients live in ONE flat fp32 arena (``params.GradArena``) laid out as three buckets in the order
  their last gradient is produced -- decoder, audio encoder, image encoder (35.5 / 5 / 5 MB; the optional cross_attn mixer last) -- so that a bucket is a contiguous slice and zeroing is one memset;
* the encoder / decoder autograd nodes call ``_bwd_done_cb`` when their last weight-gradient kernel has
  been enqueued; ``BucketReducer.mark_ready`` then launches that bucket's all-reduce immediately (NCCL
  runs it on its own stream after an event on the compute stream), i.e. the decoder bucket -- 75 % of the
  bytes -- overlaps the whole encoder backward (60 % of the backward time);
* collectives are always ISSUED in bucket order on every rank, whatever order the buckets became
  ready in, and buckets that never became ready (an encoder that received no gradient because its
  modality was dropped by ``apply_teacher_forcing_modality``, reference model.py:561-575) are reduced
  as zeros at ``finish()`` -- ranks draw the modality independently, so they must not disagree on the
  collective sequence;
* the mean over ranks is folded into the fused Adam kernel (``grad_scale = 1 / world``), so there is no
  separate scaling pass;
* InstanceNorm / LayerNorm are per-sample, so no statistic is synchronised, and inference shards the
  batch with no communication at all.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import torch
import torch.distributed as dist
import torch.nn as nn

from .params import GradArena


class BucketReducer:
    """Ordered, overlapped all-reduce (sum) of a fixed list of flat buckets.  Device-agnostic: the CPU
    ``gloo`` tests drive it with plain tensors, the trainer with slices of the gradient arena."""

    def __init__(self, buckets: Sequence[torch.Tensor], group: Optional[dist.ProcessGroup] = None):
        self.buckets = list(buckets)
        self.group = group
        self._ready = [False] * len(self.buckets)
        self._events: List = [None] * len(self.buckets)
        self._issued = 0
        self._works: List = []

    def _issue_ready_prefix(self, force: bool = False) -> None:
        while self._issued < len(self.buckets) and (force or self._ready[self._issued]):
            b = self.buckets[self._issued]
            ev = self._events[self._issued]
            if ev is not None:
                # the bucket may have been completed on ANOTHER stream than the one issuing it now (the two encoders
                # run their backward passes on two streams): order the reduction after the completing kernels
                torch.cuda.current_stream(b.device).wait_event(ev)
            self._works.append(dist.all_reduce(b, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            self._issued += 1

    def mark_ready(self, index: int) -> None:
        """bucket ``index`` has received its last gradient on the current stream"""
        self._ready[index] = True
        if self.buckets[index].is_cuda:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.buckets[index].device))
            self._events[index] = ev
        self._issue_ready_prefix()

    def finish(self) -> None:
        """issue whatever is left (in order) and make the current stream wait for every reduction"""
        self._issue_ready_prefix(force=True)
        for w in self._works:
            w.wait()
        self._works.clear()
        self._ready = [False] * len(self.buckets)
        self._events = [None] * len(self.buckets)
        self._issued = 0


def token_weight(targets: torch.Tensor, ignore_index: int = 0, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """Scale that makes data-parallel training match ONE process on the concatenated batch.

    The reference's loss is the mean over the LOCAL non-pad targets (``CrossEntropyLoss(ignore_index=pad)``, reference
    model.py:444,588), so averaging gradients over ranks (stock DDP, and this module's default) weights every rank
    equally even when their token counts differ.  The global-batch mean is ``sum_r n_r * loss_r / sum_r n_r``; with the
    gradient MEAN over ranks that is ``loss_r * s_r`` with ``s_r = n_r * world / sum_r n_r``, which this returns as a
    0-dim tensor on the targets' device (one scalar all-reduce, no host synchronisation):

        loss = model.decoder.loss(...)                        # local mean, as the reference computes it
        (loss * ddp.token_weight(y_out, pad)).backward()      # global-batch semantics

    Without an initialised process group (or world size 1) the scale is exactly 1."""
    n = (targets != ignore_index).sum().to(torch.float32)
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return torch.ones_like(n)
    total = n.clone()
    dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
    return n * dist.get_world_size(group) / total.clamp_min(1.0)


class DataParallel:
    """Wraps a ``Transformer`` / ``MultimodalTransformer`` for data-parallel training.

    usage per step::

        dp.zero_grad()
        loss = model.training_step(batch, i)      # or model.decoder.loss(...)
        loss.backward()                            # bucket all-reduces start during the backward
        dp.sync_gradients()                        # current stream waits for them
        optimizer.step()                           # FusedAdam with grad_scale = 1 / world
    """

    def __init__(self, model: nn.Module, group: Optional[dist.ProcessGroup] = None, broadcast: bool = True):
        self.model = model
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        groups = self._bucket_modules(model)
        params: List[nn.Parameter] = []
        bounds = [0]
        self._modules = []
        for mods in groups:
            for m in mods:
                params += [p for p in m.parameters() if p.requires_grad]
            bounds.append(len(params))
            self._modules.append(mods)
        self.arena = GradArena(params)
        self.buckets = [self.arena.flat[self.arena.offsets[a]: self.arena.offsets[b]] for a, b in zip(bounds[:-1], bounds[1:])]
        self.reducer = BucketReducer(self.buckets, group) if self.world > 1 else None
        for i, mods in enumerate(self._modules):
            mods[0]._bwd_done_cb = self._make_cb(i)
        if broadcast and self.world > 1:
            with torch.no_grad():
                for t in list(model.parameters()) + list(model.buffers()):
                    dist.broadcast(t, src=0, group=group)

    @staticmethod
    def _bucket_modules(model: nn.Module) -> List[List[nn.Module]]:
        # the optional cross_attn mixer runs twice per step for "attn_both" and sits between decoder and
        # encoders in the backward; it has no ready callback and goes last so that it never blocks the others
        tail = [[model.cross_attn]] if hasattr(model, "cross_attn") else []
        if hasattr(model, "image_encoder"):
            return [[model.decoder], [model.audio_encoder], [model.image_encoder]] + tail
        return [[model.decoder], [model.encoder]] + tail

    def _make_cb(self, index: int) -> Callable[[], None]:
        def cb() -> None:
            if self.reducer is not None:
                self.reducer.mark_ready(index)

        return cb

    @property
    def grad_scale(self) -> float:
        return 1.0 / self.world

    def zero_grad(self) -> None:
        if not self.arena.attached():
            self.arena.reattach()
        self.arena.zero_()

    def token_weight(self, targets: torch.Tensor, ignore_index: int = 0) -> torch.Tensor:
        """see :func:`token_weight` (global-batch loss semantics; bench.py times the stock-DDP mean of local means)"""
        return token_weight(targets, ignore_index, self.group)

    def sync_gradients(self) -> None:
        if self.reducer is not None:
            self.reducer.finish()
