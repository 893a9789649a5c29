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
                 optimizer, variants: int = 2, warmup: int = 2, double_buffer: bool = False):
        """step_fn(batch) must run ONE full step (zero grads .. optimizer.step) on the current stream, return the loss
        tensor, and must not synchronise with the host.

        double_buffer (opt-in): every captured variant owns its OWN static inputs, so that ``prefetch(batch)`` can copy
        the next batch (on a copy stream) while the current step is still running on the other variant's inputs."""
        self.step_fn = step_fn
        self.opt = optimizer
        dev = example_batch[0].device
        self.static_in: List[torch.Tensor] = [torch.empty_like(t, device=dev) for t in example_batch]
        for s, t in zip(self.static_in, example_batch):
            s.copy_(t)
        self.graphs: List[torch.cuda.CUDAGraph] = []
        self.losses: List[torch.Tensor] = []
        self._next = 0
        nvar = max(1, variants)
        self.double_buffer = bool(double_buffer) and nvar > 1
        # inputs[i] = what variant i reads: one shared set by default, a private copy per variant when double-buffered
        self.inputs: List[List[torch.Tensor]] = [self.static_in] + [
            ([t.clone() for t in self.static_in] if self.double_buffer else self.static_in) for _ in range(nvar - 1)]
        self._copy_stream = torch.cuda.Stream(device=dev) if self.double_buffer else None
        self._ready: List = [None] * nvar  # copy-stream event: variant i's inputs hold the prefetched batch
        self._done: List = [None] * nvar   # compute-stream event: variant i's last replay has consumed its inputs
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
            for v in range(nvar):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=pool, stream=cap_stream):
                    loss = self.step_fn(self.inputs[v])
                pool = g.pool()
                self.graphs.append(g)
                self.losses.append(loss)
        finally:
            ops.SEED_OFFSET_DEV = prev

    def load(self, batch: Sequence[torch.Tensor]) -> None:
        """copy ``batch`` (pinned host or device tensors) into the inputs of the variant that runs next, on the current stream"""
        for s, t in zip(self.inputs[self._next], batch):
            s.copy_(t, non_blocking=True)

    def prefetch(self, batch: Sequence[torch.Tensor]) -> None:
        """double-buffered mode: copy the batch of the NEXT call on the copy stream, overlapping whatever the compute stream
        is running; the next ``__call__()`` (without a batch) waits for the copy.  The copy itself waits until the previous
        replay of that variant has finished reading its inputs."""
        if not self.double_buffer:
            raise RuntimeError("prefetch needs GraphedTrainStep(..., double_buffer=True) and at least two variants")
        j = self._next
        cs = self._copy_stream
        if self._done[j] is not None:
            cs.wait_event(self._done[j])
        with torch.cuda.stream(cs):
            for s, t in zip(self.inputs[j], batch):
                s.copy_(t, non_blocking=True)
            self._ready[j] = cs.record_event()

    def __call__(self, batch: Optional[Sequence[torch.Tensor]] = None) -> torch.Tensor:
        """replay one step (after copying ``batch`` into the static inputs when given); returns the loss tensor of the
        replayed graph (device scalar, overwritten by the next replay of the same variant)"""
        if batch is not None:
            self.load(batch)
        i = self._next
        self._next = (i + 1) % len(self.graphs)
        if self._ready[i] is not None:
            torch.cuda.current_stream(self.inputs[i][0].device).wait_event(self._ready[i])
            self._ready[i] = None
        self.graphs[i].replay()
        if self.double_buffer:
            self._done[i] = torch.cuda.current_stream(self.inputs[i][0].device).record_event()
        return self.losses[i]
