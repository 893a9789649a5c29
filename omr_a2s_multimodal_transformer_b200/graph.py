"""Whole-training-step CUDA graphs.

One training step of the hot path is ~900 kernel launches driven from Python (custom autograd nodes composed of
C-ABI calls); at B200 speeds the host needs longer to *enqueue* them (~35 ms) than the GPU needs to run them, so the
step is captured once -- forward, loss, backward, bucketed NCCL all-reduces, fused Adam -- and replayed:

* inputs live in static device buffers (``load`` copies a batch into them, from pinned host memory or the device);
* everything random that the captured Python code draws on the HOST is frozen into a graph: the MixDropout position
  and kind of every encoder block and the dropout seeds (reference encoder.py:87-104,160).  To keep training
  stochastic (a) every dropout kernel mixes a DEVICE counter -- the optimizer's step counter -- into its seed, so
  each replay draws fresh masks, and (b) ``variants`` graphs (default 8) are captured with independent host draws and
  one of them is picked AT RANDOM (host RNG) for every replay, so the sequence of (position, kind) draws is random with
  replacement over the captured set rather than the reference's fresh draw per step -- the documented approximation of
  this fast path; the eager ``training_step`` redraws everything like the reference.  Teacher-forcing noise uses
  torch's graph-safe CUDA generator;
* host-side decisions that change the step's STRUCTURE (the teacher-forcing modality draw of reference
  model.py:561-575: "both" / "image" / "audio") are ``modes``: ``step_fn(batch, mode)`` is captured once per
  (mode, variant) and the caller draws the mode per replay on the host, exactly as the reference does;
* kernels read the Adam step from the device, so bias correction advances with every replay;
* NCCL collectives issued through ``torch.distributed`` are captured like any other stream work.

The eager path (``model.training_step`` under Lightning) is unchanged; this is the fast path for static shapes.
"""
from __future__ import annotations

import os
import random
from typing import Callable, List, Optional, Sequence

import torch

from . import ops


class GraphedTrainStep:
    def __init__(self, step_fn: Callable[..., torch.Tensor], example_batch: Sequence[torch.Tensor],
                 optimizer, variants: int = 8, warmup: int = 2, double_buffer: bool = False,
                 modes: Optional[Sequence] = None, mode_variants: Optional[dict] = None):
        """step_fn(batch) -- or step_fn(batch, mode) when ``modes`` is given -- must run ONE full step (zero grads ..
        optimizer.step) on the current stream, return the loss tensor, and must not synchronise with the host.

        variants: graphs captured per mode with independent host-side random draws (``mode_variants`` overrides the
        count for individual modes); double_buffer: TWO static input sets, used alternately, so that
        ``prefetch(batch)`` can copy the next batch (on a copy stream) while the current step still reads the other
        set; every (mode, variant) is then captured once per input set."""
        self.step_fn = step_fn
        self.opt = optimizer
        dev = example_batch[0].device
        self.static_in: List[torch.Tensor] = [torch.empty_like(t, device=dev) for t in example_batch]
        for s, t in zip(self.static_in, example_batch):
            s.copy_(t)
        self.modes = list(modes) if modes is not None else [None]
        self._next = 0  # input set of the next call
        self.double_buffer = bool(double_buffer)
        nset = 2 if self.double_buffer else 1
        # inputs[j] = static input set j
        self.inputs: List[List[torch.Tensor]] = [self.static_in] + [[t.clone() for t in self.static_in] for _ in range(nset - 1)]
        self._copy_stream = torch.cuda.Stream(device=dev) if self.double_buffer else None
        self._ready: List = [None] * nset  # copy-stream event: input set j holds the prefetched batch
        self._done: List = [None] * nset   # compute-stream event: the last replay that read input set j has finished
        self._rng = random.Random(0x5EED)
        # graphs[mode][input set] = [(graph, loss tensor) per variant]
        self.graphs: dict = {}
        # warm-up on a side stream: builds weight caches / Adam state / kernel attributes, settles the allocator
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                for mode in self.modes:
                    self._run(self.static_in, mode)
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
            for mode in self.modes:
                nvar = max(1, int((mode_variants or {}).get(mode, variants)))
                per_set = []
                for j in range(nset):
                    caps = []
                    for _ in range(nvar):
                        g = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(g, pool=pool, stream=cap_stream):
                            loss = self._run(self.inputs[j], mode)
                        pool = g.pool()
                        caps.append((g, loss))
                    per_set.append(caps)
                self.graphs[mode] = per_set
        finally:
            ops.SEED_OFFSET_DEV = prev

    def _run(self, batch, mode):
        return self.step_fn(batch) if mode is None and self.modes == [None] else self.step_fn(batch, mode)

    def num_graphs(self) -> int:
        return sum(len(caps) for per_set in self.graphs.values() for caps in per_set)

    def release(self) -> None:
        """drop the captured graphs (they hold the NCCL kernels of the gradient all-reduce: release them before
        ``destroy_process_group()``)"""
        self.graphs.clear()

    def load(self, batch: Sequence[torch.Tensor]) -> None:
        """copy ``batch`` (pinned host or device tensors) into the input set the next call reads, on the current stream"""
        for s, t in zip(self.inputs[self._next], batch):
            s.copy_(t, non_blocking=True)

    def prefetch(self, batch: Sequence[torch.Tensor]) -> None:
        """double-buffered mode: copy the batch of the NEXT call on the copy stream, overlapping whatever the compute stream
        is running; the next ``__call__()`` (without a batch) waits for the copy.  The copy itself waits until the previous
        replay of that variant has finished reading its inputs."""
        if not self.double_buffer:
            raise RuntimeError("prefetch needs GraphedTrainStep(..., double_buffer=True)")
        j = self._next
        cs = self._copy_stream
        if self._done[j] is not None:
            cs.wait_event(self._done[j])
        with torch.cuda.stream(cs):
            for s, t in zip(self.inputs[j], batch):
                s.copy_(t, non_blocking=True)
            self._ready[j] = cs.record_event()

    def __call__(self, batch: Optional[Sequence[torch.Tensor]] = None, mode=None) -> torch.Tensor:
        """replay one step (after copying ``batch`` into the static inputs when given) of ``mode`` (default: the first
        mode); the variant is drawn at random.  Returns the loss tensor of the replayed graph (device scalar, overwritten
        by the next replay of the same graph)"""
        if batch is not None:
            self.load(batch)
        j = self._next
        self._next = (j + 1) % len(self.inputs)
        if self._ready[j] is not None:
            torch.cuda.current_stream(self.inputs[j][0].device).wait_event(self._ready[j])
            self._ready[j] = None
        caps = self.graphs[self.modes[0] if mode is None else mode][j]
        g, loss = caps[self._rng.randrange(len(caps))] if len(caps) > 1 else caps[0]
        g.replay()
        if self.double_buffer:
            self._done[j] = torch.cuda.current_stream(self.inputs[j][0].device).record_event()
        return loss
