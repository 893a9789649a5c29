"""Whole-training-step CUDA graphs.

One training step of the hot path is ~900 kernel launches driven from Python (custom autograd nodes composed of
C-ABI calls); at B200 speeds the host needs longer to *enqueue* them (~35 ms) than the GPU needs to run them, so the
step is captured once -- forward, loss, backward, bucketed NCCL all-reduces, fused Adam -- and replayed:

* inputs live in static device buffers (``load`` copies a batch into them, from pinned host memory or the device);
* everything random that the captured Python code draws on the HOST is frozen into a graph: the MixDropout position
  and kind of every encoder block and the dropout seeds (reference encoder.py:87-104,160).  To keep training
  stochastic (a) every dropout kernel mixes a DEVICE counter -- the optimizer's step counter -- into its seed, so
  each replay draws fresh masks, and (b) ``variants`` graphs are captured with independent host draws and replayed
  round-robin; teacher-forcing noise uses torch's graph-safe CUDA generator;
* kernels read the Adam step from the device, so bias correction advances with every replay;
* NCCL collectives issued through ``torch.distributed`` are captured like any other stream work.

The eager path (``model.training_step`` under Lightning) is unchanged; this is the fast path for static shapes.
"""
from __future__ import annotations

import os
from typing import Callable, List, Optional, Sequence

import torch

from . import ops


class GraphedTrainStep:
    def __init__(self, step_fn: Callable[[Sequence[torch.Tensor]], torch.Tensor], example_batch: Sequence[torch.Tensor],
                 optimizer, variants: int = 2, warmup: int = 2):
        """step_fn(batch) must run ONE full step (zero grads .. optimizer.step) on the current stream, return the loss
        tensor, and must not synchronise with the host."""
        self.step_fn = step_fn
        self.opt = optimizer
        dev = example_batch[0].device
        self.static_in: List[torch.Tensor] = [torch.empty_like(t, device=dev) for t in example_batch]
        for s, t in zip(self.static_in, example_batch):
            s.copy_(t)
        self.graphs: List[torch.cuda.CUDAGraph] = []
        self.losses: List[torch.Tensor] = []
        self._next = 0
        # warm-up on a side stream: builds weight caches / Adam state / kernel attributes, settles the allocator
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                self.step_fn(self.static_in)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        step_dev = getattr(optimizer, "_step_dev", None)
        if step_dev is None:
            raise RuntimeError("GraphedTrainStep needs the FusedAdam device step counter (run at least one optimizer step)")
        prev = ops.SEED_OFFSET_DEV
        ops.SEED_OFFSET_DEV = step_dev
        # The dependent chains (encoders' data path, decoder chain) are captured on a HIGH-priority stream; the off-chain
        # work forked from them (weight gradients, cross-K/V projections: decoder._side_stream, encoder._wgrad_stream)
        # runs on default-priority streams, so whenever both have CTAs pending the chain goes first
        # (OMR_STREAM_PRIORITY=0: capture on a default-priority stream).
        prio = os.environ.get("OMR_STREAM_PRIORITY", "1") != "0"
        cap_stream = torch.cuda.Stream(device=dev, priority=-1) if prio else torch.cuda.Stream(device=dev)
        try:
            pool = None
            for _ in range(max(1, variants)):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=pool, stream=cap_stream):
                    loss = self.step_fn(self.static_in)
                pool = g.pool()
                self.graphs.append(g)
                self.losses.append(loss)
        finally:
            ops.SEED_OFFSET_DEV = prev

    def load(self, batch: Sequence[torch.Tensor]) -> None:
        for s, t in zip(self.static_in, batch):
            s.copy_(t, non_blocking=True)

    def __call__(self, batch: Optional[Sequence[torch.Tensor]] = None) -> torch.Tensor:
        """replay one step (after copying ``batch`` into the static inputs when given); returns the loss tensor of the
        replayed graph (device scalar, overwritten by the next replay of the same variant)"""
        if batch is not None:
            self.load(batch)
        i = self._next
        self._next = (i + 1) % len(self.graphs)
        self.graphs[i].replay()
        return self.losses[i]
