"""Parameter-side plumbing shared by the encoder/decoder modules.

* ``WeightCache`` keeps the kernel-layout working copies of the fp32 ``nn.Parameter``s (bf16 or fp32,
  re-laid-out for the implicit-GEMM kernels) and refreshes them when a parameter's version changes.
* ``grad_buf`` returns the fp32 gradient buffer of a parameter that the weight-gradient kernels
  accumulate into directly (``param.grad``, created zero-filled on first use, or a slice of a
  ``GradArena``).
* ``GradArena`` lays all gradients of a model out in one flat fp32 buffer so that zeroing is one
  memset and the data-parallel all-reduce can run on contiguous buckets.
* ``resolve_dtype`` picks the kernel storage type (fp32 exact mode / bf16 tensor-core mode).
"""
from __future__ import annotations

import os
from typing import Dict, Iterable, List, Optional, Tuple

import torch
import torch.nn as nn

from . import ops


def resolve_dtype(explicit: Optional[torch.dtype]) -> torch.dtype:
    """Explicit setting wins; else env OMR_COMPUTE_DTYPE; else bf16 under torch autocast (what
    Lightning's ``precision="16-mixed"`` of the reference's train.py:153 enables), else fp32."""
    if explicit is not None:
        return explicit
    env = os.environ.get("OMR_COMPUTE_DTYPE", "").lower()
    if env in ("bf16", "bfloat16"):
        return torch.bfloat16
    if env in ("fp32", "float32"):
        return torch.float32
    if torch.is_autocast_enabled():
        return torch.bfloat16
    return torch.float32


class WeightCache:
    """kind: "conv" [Co,3,3,Ci] | "convT" [Ci,3,3,Co] | "dw" [3,3,C] | "mat" [N,K] (same order, 2-D) | "matT" [K,N] |
    "matKS4" [4][N][K/4] | "matDecA" / "matDecKS" (mma fragment orders of the bf16 persistent decode kernel) -- the last
    three are not Adam shadows: re-packed when the parameter's version moves."""

    def __init__(self) -> None:
        self._c: Dict[Tuple[int, str, torch.dtype], Tuple[int, torch.Tensor]] = {}

    def clear(self) -> None:
        self._c.clear()

    def get(self, p: torch.Tensor, kind: str, dtype: torch.dtype) -> torch.Tensor:
        key = (id(p), kind, dtype)
        ent = self._c.get(key)
        ver = p._version
        if ent is not None and ent[0] == ver and ent[1].device == p.device and ent[2] == p.data_ptr():
            return ent[1]
        t = self._pack(p.detach(), kind, dtype)
        self._c[key] = (ver, t, p.data_ptr())
        return t

    def peek(self, p: torch.Tensor, kind: str, dtype: torch.dtype) -> Optional[torch.Tensor]:
        ent = self._c.get((id(p), kind, dtype))
        return None if ent is None else ent[1]

    KIND_LAYOUT = {"mat": 0, "conv": 1, "dw": 2, "convT": 3, "matT": 4}  # omr_adam_entry.layout codes

    def shadows(self, p: torch.Tensor):
        """bf16 working copies of ``p`` currently cached: [(tensor, adam layout code)] (for FusedAdam, which
        rewrites them in the same pass as the fp32 master so they never need re-packing)."""
        out = []
        for kind, code in self.KIND_LAYOUT.items():
            ent = self._c.get((id(p), kind, torch.bfloat16))
            if ent is not None and ent[1].device == p.device and ent[2] == p.data_ptr():
                out.append((ent[1], code))
        return out

    def mark_fresh(self, p: torch.Tensor) -> None:
        """the optimizer kernel has just rewritten every bf16 copy of ``p`` returned by ``shadows``"""
        ver = p._version
        for kind in self.KIND_LAYOUT:
            key = (id(p), kind, torch.bfloat16)
            ent = self._c.get(key)
            if ent is not None:
                self._c[key] = (ver, ent[1], ent[2])

    @staticmethod
    def _pack(w: torch.Tensor, kind: str, dtype: torch.dtype) -> torch.Tensor:
        if w.dtype != torch.float32:
            raise TypeError("parameters must stay float32 (the kernels keep bf16 working copies themselves)")
        w = w.contiguous()
        if kind == "conv":
            return ops.pack_conv_weight(w, dtype, transpose=False)
        if kind == "convT":
            return ops.pack_conv_weight(w, dtype, transpose=True)
        if kind == "dw":
            return ops.pack_dw_weight(w, dtype)
        if kind == "mat":
            w2 = w.reshape(w.shape[0], -1)
            return w2 if dtype == torch.float32 else ops.cast(w2, dtype)
        if kind == "matKS4":  # [N,K] -> [4][N][K/4]: the column slices of the persistent decode kernel's K-split projections
            w2 = w.reshape(w.shape[0], -1)
            n, k = w2.shape
            return ops.cast(w2.view(n, 4, k // 4).permute(1, 0, 2).contiguous(), dtype).view(4 * n, k // 4)
        if kind == "matDecA":
            # [N,K] -> mma.sync A-fragment order of the persistent decode kernel's bf16 projections: rows padded to a
            # multiple of 32, then [N/16 m-tiles][K/32 k-blocks][2 row halves][8 rows g][4 lanes t][8 elements]: a warp's
            # LDS.128 of one (m-tile, k-block, half) is 512 contiguous bytes, lane (g,t) gets row g (+8), k = 32kb+8t..+7
            w2 = w.reshape(w.shape[0], -1)
            n, k = w2.shape
            npad = (n + 31) // 32 * 32
            if npad != n:
                w2 = torch.cat([w2, w2.new_zeros(npad - n, k)], 0)
            t = w2.view(npad // 16, 2, 8, k // 32, 4, 8).permute(0, 3, 1, 2, 4, 5).contiguous()
            return ops.cast(t, dtype).view(npad, k)
        if kind == "matDecKS":
            # [N,K] -> the four column slices [4][N][K/4] of "matKS4", each in the A-fragment order above
            w2 = w.reshape(w.shape[0], -1)
            n, k = w2.shape
            ks = k // 4
            t = w2.view(n // 16, 2, 8, 4, ks // 32, 4, 8).permute(3, 0, 4, 1, 2, 5, 6).contiguous()
            return ops.cast(t, dtype).view(4 * n, ks)
        if kind == "matT":  # [R,C] -> [C,R]: the K-major operand of the data-gradient GEMMs
            w2 = w.reshape(w.shape[0], -1)
            return ops.cast(w2.t().contiguous(), dtype)
        raise ValueError(kind)


def grad_buf(p: nn.Parameter) -> torch.Tensor:
    if p.grad is None:
        p.grad = torch.zeros_like(p, memory_format=torch.contiguous_format)
    return p.grad


class GradArena:
    """All gradients of ``params`` as views of one flat fp32 buffer (in the given order)."""

    def __init__(self, params: Iterable[nn.Parameter]) -> None:
        self.params: List[nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("GradArena: no trainable parameters")
        dev = self.params[0].device
        sizes = [(p.numel() + 63) // 64 * 64 for p in self.params]  # 256-byte aligned slots
        self.offsets = [0]
        for s in sizes:
            self.offsets.append(self.offsets[-1] + s)
        self.flat = torch.zeros(self.offsets[-1], dtype=torch.float32, device=dev)
        for p, o in zip(self.params, self.offsets):
            p.grad = self.flat[o : o + p.numel()].view_as(p)

    def zero_(self) -> None:
        self.flat.zero_()

    def attached(self) -> bool:
        return all(p.grad is not None and p.grad.data_ptr() == self.flat.data_ptr() + 4 * o
                   for p, o in zip(self.params, self.offsets))

    def reattach(self) -> None:
        for p, o in zip(self.params, self.offsets):
            p.grad = self.flat[o : o + p.numel()].view_as(p)


# ---- parameter containers with torch-compatible names and default initialisation ---------------------
class ConvParams(nn.Module):
    """weight/bias of an nn.Conv2d / nn.Conv1d call site (same shapes, same default init)."""

    def __init__(self, in_c: int, out_c: int, ksize: Tuple[int, ...], groups: int = 1) -> None:
        super().__init__()
        self.in_channels, self.out_channels, self.kernel_size, self.groups = in_c, out_c, tuple(ksize), groups
        self.weight = nn.Parameter(torch.empty(out_c, in_c // groups, *ksize))
        self.bias = nn.Parameter(torch.empty(out_c))
        self.reset_parameters()

    def reset_parameters(self) -> None:
        nn.init.kaiming_uniform_(self.weight, a=5 ** 0.5)
        fan_in = self.weight[0].numel()
        bound = 1.0 / fan_in ** 0.5 if fan_in > 0 else 0.0
        nn.init.uniform_(self.bias, -bound, bound)


class LinearParams(nn.Module):
    def __init__(self, in_f: int, out_f: int, zero_bias: bool = False) -> None:
        super().__init__()
        self.in_features, self.out_features = in_f, out_f
        self.weight = nn.Parameter(torch.empty(out_f, in_f))
        self.bias = nn.Parameter(torch.empty(out_f))
        nn.init.kaiming_uniform_(self.weight, a=5 ** 0.5)
        if zero_bias:
            nn.init.zeros_(self.bias)
        else:
            bound = 1.0 / in_f ** 0.5
            nn.init.uniform_(self.bias, -bound, bound)


class LayerNormParams(nn.Module):
    def __init__(self, d: int, eps: float = 1e-5) -> None:
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(d))
        self.bias = nn.Parameter(torch.zeros(d))


class MHAParams(nn.Module):
    """Parameters of an nn.MultiheadAttention (packed in-proj): same names, same default init."""

    def __init__(self, embed_dim: int, num_heads: int) -> None:
        super().__init__()
        self.embed_dim, self.num_heads, self.head_dim = embed_dim, num_heads, embed_dim // num_heads
        self.in_proj_weight = nn.Parameter(torch.empty(3 * embed_dim, embed_dim))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * embed_dim))
        self.out_proj = LinearParams(embed_dim, embed_dim, zero_bias=True)
        nn.init.xavier_uniform_(self.in_proj_weight)
