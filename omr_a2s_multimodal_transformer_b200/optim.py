"""Fused multi-tensor Adam (``torch.optim.Adam(lr=1e-4, betas=(0.9, 0.999), eps=1e-8, amsgrad=False)`` of
reference model.py:134-139,475-483) as ONE kernel launch over all parameters.

The kernel reads fp32 gradients, updates the fp32 master parameters and both moments, and -- when the
model runs in bf16 -- refreshes the bf16 kernel-layout working copies of the weights in the same pass, so
no separate cast/re-layout kernels run after the step.  The step counter lives on the device, which makes
the whole optimizer step CUDA-graph capturable.
"""
from __future__ import annotations

import ctypes
from typing import Iterable, List, Optional

import torch

from . import _lib
from ._lib import call, ptr, stream_ptr


class _AdamEntry(ctypes.Structure):
    _fields_ = [
        ("param", ctypes.c_void_p), ("grad", ctypes.c_void_p), ("exp_avg", ctypes.c_void_p),
        ("exp_avg_sq", ctypes.c_void_p), ("shadow", ctypes.c_void_p), ("shadow2", ctypes.c_void_p),
        ("n", ctypes.c_longlong), ("layout", ctypes.c_int), ("layout2", ctypes.c_int), ("d0", ctypes.c_int),
        ("d1", ctypes.c_int),
    ]


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                 grad_scale: float = 1.0):
        super().__init__(list(params), dict(lr=lr, betas=betas, eps=eps))
        self.grad_scale = grad_scale
        self._table = None
        self._table_key = None
        self._step_dev: Optional[torch.Tensor] = None
        self._shadows = []  # callables returning [(param, shadow, layout, shadow2, layout2, d0, d1)]

    def register_shadow_provider(self, fn, mark_fresh=None) -> None:
        """fn(param) -> [(bf16 tensor, layout code)] (at most two are kept in sync by the kernel; any further
        copy is simply re-packed by its WeightCache); mark_fresh(param) is called after every step for the
        parameters whose copies the kernel rewrote."""
        self._shadows.append((fn, mark_fresh))
        self._table = None

    def _params(self) -> List[torch.nn.Parameter]:
        return [p for g in self.param_groups for p in g["params"]]

    def _build_table(self) -> None:
        ps = [p for p in self._params() if p.requires_grad]
        dev = ps[0].device
        if self._step_dev is None or self._step_dev.device != dev:
            self._step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        entries = (_AdamEntry * len(ps))()
        key = []
        self._has_shadow = []
        for i, p in enumerate(ps):
            st = self.state[p]
            if "exp_avg" not in st:
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
            sh = self._query_shadows(p)
            e = entries[i]
            e.param, e.grad = p.data_ptr(), (p.grad.data_ptr() if p.grad is not None else None)
            e.exp_avg, e.exp_avg_sq = st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
            e.n = p.numel()
            if len(sh) > 2:
                sh = []  # more copies than the kernel keeps: let the caches re-pack all of them
            if sh:
                e.shadow, e.layout = sh[0][0].data_ptr(), sh[0][1]
                if len(sh) > 1:
                    e.shadow2, e.layout2 = sh[1][0].data_ptr(), sh[1][1]
                e.d0 = p.shape[0]
                e.d1 = p.shape[1] if p.dim() > 1 else 1
                if p.dim() == 2 or (p.dim() == 3 and p.shape[2] == 1):
                    e.d1 = p.shape[1]
            key.append((e.param, e.grad, tuple(t.data_ptr() for t, _ in sh), e.exp_avg, e.exp_avg_sq))
            self._has_shadow.append(bool(sh))
        raw = bytes(entries)
        host = torch.frombuffer(bytearray(raw), dtype=torch.uint8)
        self._table = host.to(dev)
        self._table_key = key
        self._n = len(ps)
        self._max_n = max(p.numel() for p in ps)
        self._ps = ps

    def _query_shadows(self, p):
        out = []
        for fn, _ in self._shadows:
            out += fn(p) or []
        return out

    def _table_stale(self) -> bool:
        if self._table is None:
            return True
        for p, k in zip(self._ps, self._table_key):
            g = p.grad.data_ptr() if p.grad is not None else None
            if p.data_ptr() != k[0] or g != k[1]:
                return True
            sh = self._query_shadows(p) if self._shadows else []
            if len(sh) > 2:
                sh = []
            if tuple(t.data_ptr() for t, _ in sh) != k[2]:
                return True
            st = self.state.get(p)
            if st is None or "exp_avg" not in st or st["exp_avg"].data_ptr() != k[3] or st["exp_avg_sq"].data_ptr() != k[4]:
                return True  # load_state_dict() replaced the moment tensors: the kernel must not keep the old pointers
        return False

    # ---- checkpointing: same layout as torch.optim.Adam (per-parameter ``step``, ``exp_avg``, ``exp_avg_sq``) -------------
    def state_dict(self):
        """The Adam step lives on the device (``_step_dev``, shared by all parameters); it is written into every
        parameter's ``state["step"]`` here so that a checkpoint carries it exactly like torch.optim.Adam's (one host
        read at checkpoint time)."""
        if self._step_dev is not None:
            t = float(self._step_dev.item())
            for p in self._params():
                st = self.state.get(p)
                if st is not None and "exp_avg" in st:
                    st["step"] = torch.tensor(t, dtype=torch.float32)
        return super().state_dict()

    def load_state_dict(self, state_dict) -> None:
        """Restores the moments AND the bias-correction step (from this class or from a torch.optim.Adam checkpoint), and
        drops the pointer table so that the next step() binds the loaded moment tensors."""
        super().load_state_dict(state_dict)
        steps = [float(st["step"]) for st in self.state.values() if "step" in st]
        ps = [p for p in self._params() if p.requires_grad]
        if steps and ps:
            if max(steps) != min(steps):
                raise RuntimeError("FusedAdam keeps ONE step counter for all parameters; the checkpoint holds several "
                                   f"({min(steps)} .. {max(steps)})")
            dev = ps[0].device
            if self._step_dev is None or self._step_dev.device != dev:
                self._step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
            self._step_dev.fill_(int(steps[0]))
        for p in ps:  # moments must be contiguous fp32 on the parameter's device (the kernel reads raw pointers)
            st = self.state.get(p)
            if st is not None and "exp_avg" in st:
                for k in ("exp_avg", "exp_avg_sq"):
                    st[k] = st[k].to(device=p.device, dtype=p.dtype).contiguous()
        self._table = None

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        if self._table_stale():
            self._build_table()
        g = self.param_groups[0]
        call("omr_adam_tick", ptr(self._step_dev), stream_ptr())
        call("omr_adam_step", ptr(self._table), self._n, self._max_n, ptr(self._step_dev), float(g["lr"]),
             float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]), float(self.grad_scale), stream_ptr())
        # the kernel wrote the parameters through raw pointers: bump their autograd version so that every
        # cached re-layout (params.WeightCache) is refreshed, except the bf16 copies the kernel itself rewrote
        torch.autograd.graph.increment_version(self._ps)
        for p, has in zip(self._ps, self._has_shadow):
            if has:
                for _, mark in self._shadows:
                    if mark is not None:
                        mark(p)
        return loss
