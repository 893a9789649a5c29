"""Thin functional wrappers over the C ABI (``include/omr_b200.h``): one Python function per kernel
entry point, taking/returning CUDA tensors.  No autograd here -- the backward passes are composed
explicitly in ``encoder.py`` / ``decoder.py`` / ``model.py`` from these same wrappers.

Tensors are NHWC for image-like data and [B,T,D] for sequences; ``x.dtype`` (float32 or bfloat16)
selects the kernel storage type.  Everything is enqueued on the current torch CUDA stream.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import call, dt_code, ptr, stream_ptr

NEG_INF = float("-inf")


def _chk(t: torch.Tensor, name: str) -> torch.Tensor:
    _lib.require_cuda(t, name)
    if not t.is_contiguous():
        raise RuntimeError(f"{name}: expected a contiguous tensor, got strides {t.stride()} for shape {tuple(t.shape)}")
    return t


def out_hw(h: int, w: int, stride: Tuple[int, int]) -> Tuple[int, int]:
    return (h + stride[0] - 1) // stride[0], (w + stride[1] - 1) // stride[1]


# ---- elementwise / layout ----------------------------------------------------------------------
def cast(x: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    _chk(x, "cast")
    if x.dtype == dtype:
        return x
    y = torch.empty_like(x, dtype=dtype)
    call("omr_cast", dt_code(x.dtype), dt_code(dtype), ptr(x), ptr(y), x.numel(), stream_ptr())
    return y


def relu_bwd(y: torch.Tensor, dy: torch.Tensor, inplace: bool = False) -> torch.Tensor:
    _chk(y, "relu_bwd.y"), _chk(dy, "relu_bwd.dy")
    dx = dy if inplace else torch.empty_like(dy)
    call("omr_relu_bwd", dt_code(y.dtype), ptr(y), ptr(dy), ptr(dx), y.numel(), stream_ptr())
    return dx


def add(a: torch.Tensor, b: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _chk(a, "add.a"), _chk(b, "add.b")
    out = torch.empty_like(a) if out is None else out
    call("omr_add", dt_code(a.dtype), ptr(a), ptr(b), ptr(out), a.numel(), stream_ptr())
    return out


# Debug probe for the parity tests: when set to a list, every ReLU output of the encoders and of the decoder's feed-forward
# blocks is appended to it in call order (tests/helpers.py compares these ReLU decisions with the CPU checker's).
RELU_PROBE: Optional[list] = None


def _probe_relu(y: torch.Tensor) -> None:
    if RELU_PROBE is not None:
        RELU_PROBE.append(y)


# device int32 scalar mixed into every dropout seed on the GPU (None: host seeds only).  A graph-captured training
# step points this at its step counter so that replays draw fresh masks (graph.GraphedTrainStep).
SEED_OFFSET_DEV: Optional[torch.Tensor] = None


def dropout(x: torch.Tensor, p: float, seed: int, channelwise: bool = False, inplace: bool = False) -> torch.Tensor:
    """Seeded dropout; for NHWC tensors ``channelwise`` drops whole (sample, channel) planes (Dropout2d).
    Applying it to a gradient with the same seed is the backward pass."""
    _chk(x, "dropout")
    y = x if inplace else torch.empty_like(x)
    c = x.shape[-1]
    per_sample = x.numel() // x.shape[0]
    call("omr_dropout", dt_code(x.dtype), ptr(x), ptr(y), x.numel(), c, per_sample, float(p), int(seed), int(channelwise),
         ptr(SEED_OFFSET_DEV), stream_ptr())
    return y


def pack_conv_weight(w: torch.Tensor, dtype: torch.dtype, transpose: bool) -> torch.Tensor:
    """[Co,Ci,3,3] fp32 -> [Co,3,3,Ci] (transpose=False) or [Ci,3,3,Co] (transpose=True) in ``dtype``."""
    _chk(w, "pack_conv_weight")
    co, ci = w.shape[0], w.shape[1]
    shape = (ci, 3, 3, co) if transpose else (co, 3, 3, ci)
    out = torch.empty(shape, dtype=dtype, device=w.device)
    call("omr_pack_conv_weight", dt_code(dtype), ptr(w), ptr(out), co, ci, int(transpose), stream_ptr())
    return out


def pack_dw_weight(w: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    _chk(w, "pack_dw_weight")
    c = w.shape[0]
    out = torch.empty((3, 3, c), dtype=dtype, device=w.device)
    call("omr_pack_dw_weight", dt_code(dtype), ptr(w), ptr(out), c, stream_ptr())
    return out


# ---- convolutions ---------------------------------------------------------------------------------
def conv3x3_fwd(x, wp, bias, stride=(1, 1), relu=False, in_sums=None):
    """x [N,H,W,Ci], wp [Co,3,3,Ci] (packed), bias fp32 [Co] -> y [N,Ho,Wo,Co].
    in_sums (fp64 [N,Co,2]): receives (sum y, sum y^2) per (sample, channel) -- the statistics of the InstanceNorm that
    follows -- from the convolution's own epilogue (``instnorm_fwd(..., sums=in_sums)`` then skips its statistics pass)"""
    _chk(x, "conv3x3_fwd.x"), _chk(wp, "conv3x3_fwd.w")
    n, h, w, ci = x.shape
    co = wp.shape[0]
    ho, wo = out_hw(h, w, stride)
    y = torch.empty((n, ho, wo, co), dtype=x.dtype, device=x.device)
    call("omr_conv3x3_fwd", dt_code(x.dtype), ptr(x), ptr(wp), ptr(bias), ptr(y), n, h, w, ci, co, stride[0], stride[1],
         int(relu), ptr(in_sums), stream_ptr())
    return y


def in_sums_buffer(n: int, c: int, device) -> torch.Tensor:
    """fp64 [N,C,2] scratch for per-(sample, channel) InstanceNorm sums"""
    return torch.empty((n, c, 2), dtype=torch.float64, device=device)


def conv3x3_dgrad(dy, wpt, in_hw, stride=(1, 1), mask=None, mask_scale: float = 1.0, colsum=None, in_x=None, in_bsums=None):
    """dy [N,Ho,Wo,Co], wpt [Ci,3,3,Co] -> dx [N,H,W,Ci]; with ``mask`` (the conv's forward input, a ReLU/dropout output)
    the backward of that ReLU/dropout is fused: dx = mask > 0 ? dx * mask_scale : 0.
    colsum (fp32 [Ci]): += column sums of dx (bias gradient of the layer that produced the conv's input);
    in_x + in_bsums (fp64 [N,Ci,2]): dx is the gradient of an InstanceNorm output, in_x that norm's input: in_bsums receives
    the raw backward sums (sum dx, sum dx * in_x) for ``instnorm_bwd(..., sums=in_bsums)``"""
    _chk(dy, "conv3x3_dgrad.dy"), _chk(wpt, "conv3x3_dgrad.w")
    if mask is not None:
        _chk(mask, "conv3x3_dgrad.mask")
    if in_x is not None:
        _chk(in_x, "conv3x3_dgrad.in_x")
    n, _, _, co = dy.shape
    ci = wpt.shape[0]
    h, w = in_hw
    dx = torch.empty((n, h, w, ci), dtype=dy.dtype, device=dy.device)
    call("omr_conv3x3_dgrad", dt_code(dy.dtype), ptr(dy), ptr(wpt), ptr(dx), n, h, w, ci, co, stride[0], stride[1],
         ptr(mask), float(mask_scale), ptr(colsum), ptr(in_x), ptr(in_bsums), stream_ptr())
    return dx


def conv3x3_wgrad(x, dy, dw, db, stride=(1, 1), accumulate=True):
    """accumulates dw [Co,Ci,3,3] fp32 and db [Co] fp32 (torch parameter layouts)"""
    _chk(x, "conv3x3_wgrad.x"), _chk(dy, "conv3x3_wgrad.dy")
    n, h, w, ci = x.shape
    co = dy.shape[3]
    # scratch for the wide layers' vectorised partial-sum reduction (see omr_conv3x3_wgrad in include/omr_b200.h)
    ws = torch.empty(9 * co * ci, dtype=torch.float32, device=x.device) if (ci * co >= 4096 and x.dtype == torch.bfloat16) else None
    call("omr_conv3x3_wgrad", dt_code(x.dtype), ptr(x), ptr(dy), ptr(dw), ptr(db), n, h, w, ci, co, stride[0], stride[1],
         int(accumulate), ptr(ws), stream_ptr())


def dwconv3x3_fwd(x, wp, bias):
    _chk(x, "dwconv3x3_fwd.x")
    n, h, w, c = x.shape
    y = torch.empty_like(x)
    call("omr_dwconv3x3_fwd", dt_code(x.dtype), ptr(x), ptr(wp), ptr(bias), ptr(y), n, h, w, c, stream_ptr())
    return y


def dwconv3x3_dgrad(dy, wp):
    _chk(dy, "dwconv3x3_dgrad.dy")
    n, h, w, c = dy.shape
    dx = torch.empty_like(dy)
    call("omr_dwconv3x3_dgrad", dt_code(dy.dtype), ptr(dy), ptr(wp), ptr(dx), n, h, w, c, stream_ptr())
    return dx


def dwconv3x3_wgrad(x, dy, dw, db, accumulate=True):
    _chk(x, "dwconv3x3_wgrad.x"), _chk(dy, "dwconv3x3_wgrad.dy")
    n, h, w, c = x.shape
    call("omr_dwconv3x3_wgrad", dt_code(x.dtype), ptr(x), ptr(dy), ptr(dw), ptr(db), n, h, w, c, int(accumulate),
         stream_ptr())


def instnorm_fwd(x, eps: float, sums=None):
    """x [N,H,W,C] -> (y, stats [N,C,2] fp32 = (mean, rstd)); ``sums``: (sum x, sum x^2) already accumulated by the
    convolution that produced x (``conv3x3_fwd(..., in_sums=)``)"""
    _chk(x, "instnorm_fwd.x")
    n, h, w, c = x.shape
    y = torch.empty_like(x)
    stats = torch.empty((n, c, 2), dtype=torch.float32, device=x.device)
    ws = torch.empty((n, c, 2), dtype=torch.float64, device=x.device) if sums is None else sums
    call("omr_instnorm_fwd", dt_code(x.dtype), ptr(x), ptr(y), ptr(stats), ptr(ws), n, h * w, c, float(eps),
         int(sums is not None), stream_ptr())
    return y, stats


def instnorm_bwd(dy, x, stats, relu_mask: bool = False, mask_scale: float = 1.0, sums=None, colsum=None):
    """relu_mask: x is a ReLU (+dropout) output and the backward of that ReLU/dropout is fused into the result;
    ``sums``: raw backward sums (sum dy, sum dy * x) from ``conv3x3_dgrad(..., in_x=x, in_bsums=)``;
    colsum (fp32 [C]): += column sums of dx (bias gradient of the convolution in front of the norm)"""
    _chk(dy, "instnorm_bwd.dy"), _chk(x, "instnorm_bwd.x")
    n, h, w, c = x.shape
    dx = torch.empty_like(x)
    ws = torch.empty((n, c, 2), dtype=torch.float64, device=x.device) if sums is None else sums
    call("omr_instnorm_bwd", dt_code(x.dtype), ptr(dy), ptr(x), ptr(stats), ptr(dx), ptr(ws), n, h * w, c, int(relu_mask),
         float(mask_scale), int(sums is not None), ptr(colsum), stream_ptr())
    return dx


def pe2d_add(x, pe_nhwc, out, row_off: int):
    """out[b, row_off + p, :] = x[b, p, :] + pe[p // w, p % w, :] ; x [B,h,w,C], out [B,S,C]"""
    _chk(x, "pe2d_add.x"), _chk(out, "pe2d_add.out")
    b, h, w, c = x.shape
    if h > pe_nhwc.shape[0] or w > pe_nhwc.shape[1]:
        raise RuntimeError(f"feature map {h}x{w} exceeds the positional-encoding table {tuple(pe_nhwc.shape[:2])}")
    call("omr_pe2d_add", dt_code(x.dtype), ptr(x), ptr(pe_nhwc), ptr(out), b, h, w, c, pe_nhwc.shape[1], out.shape[1],
         row_off, stream_ptr())
    return out


def copy_rows(src, row_off: int, rows: int):
    """src [B,S,C] -> contiguous [B,rows,C] copy of src[:, row_off:row_off+rows]"""
    _chk(src, "copy_rows.src")
    b, s, c = src.shape
    dst = torch.empty((b, rows, c), dtype=src.dtype, device=src.device)
    call("omr_copy_rows", dt_code(src.dtype), ptr(src), ptr(dst), b, rows, c, s, row_off, stream_ptr())
    return dst


# ---- masks --------------------------------------------------------------------------------------
def key_bias_from_lengths(bias, lens_i32, seg_off: int, seg_len: int, value: float):
    _chk(bias, "key_bias_from_lengths.bias"), _chk(lens_i32, "key_bias_from_lengths.lens")
    assert lens_i32.dtype == torch.int32
    b, s = bias.shape
    call("omr_key_bias_from_lengths", ptr(bias), ptr(lens_i32), b, s, seg_off, seg_len, float(value), stream_ptr())
    return bias


def key_bias_from_tokens(tokens, pad_id: int, value: float):
    _chk(tokens, "key_bias_from_tokens.tokens")
    assert tokens.dtype == torch.int64
    bias = torch.empty(tokens.shape, dtype=torch.float32, device=tokens.device)
    call("omr_key_bias_from_tokens", ptr(bias), ptr(tokens), tokens.numel(), pad_id, float(value), stream_ptr())
    return bias


# ---- decoder pieces -----------------------------------------------------------------------------
def embed_pe_fwd(tokens, table, pe, pos0: int = 0, pos_dev=None, out=None):
    """tokens int64 [B,T], table [V,D] (compute dtype), pe fp32 [max_len,D] -> [B,T,D]"""
    _chk(tokens, "embed_pe_fwd.tokens"), _chk(table, "embed_pe_fwd.table")
    b, t = tokens.shape
    d = table.shape[1]
    if pos0 + t > pe.shape[0]:
        raise RuntimeError(f"sequence positions {pos0}..{pos0 + t} exceed the 1-D PE table ({pe.shape[0]})")
    out = torch.empty((b, t, d), dtype=table.dtype, device=table.device) if out is None else out
    call("omr_embed_pe_fwd", dt_code(table.dtype), ptr(tokens), ptr(table), ptr(pe), ptr(out), b, t, d, pos0, ptr(pos_dev),
         stream_ptr())
    return out


def embed_bwd(tokens, dout, dtable, padding_idx: int):
    _chk(tokens, "embed_bwd.tokens"), _chk(dout, "embed_bwd.dout")
    call("omr_embed_bwd", dt_code(dout.dtype), ptr(tokens), ptr(dout), ptr(dtable), tokens.numel(), dout.shape[-1],
         padding_idx, stream_ptr())


def gemm(a, b, c, m, n, k, *, trans_a=False, trans_b=False, lda, ldb, ldc, batch=1, stride_a=0, stride_b=0, stride_c=0,
         bias=None, bias_mode=0, relu=False, accumulate=False):
    """C[b] = act(opA(A[b]) @ opB(B[b]) + bias) (+C[b]); see omr_gemm in include/omr_b200.h"""
    call("omr_gemm", dt_code(a.dtype), dt_code(c.dtype), int(trans_a), int(trans_b), m, n, k, ptr(a), lda, stride_a,
         ptr(b), ldb, stride_b, ptr(c), ldc, stride_c, batch, ptr(bias), bias_mode if bias is not None else 0, int(relu),
         int(accumulate), stream_ptr())
    return c


def _rows2d(t: torch.Tensor, name: str) -> torch.Tensor:
    """2-D operand whose rows are contiguous (the row stride may exceed the width: column slices, padded rows)"""
    _lib.require_cuda(t, name)
    if t.dim() != 2 or (t.shape[1] > 1 and t.stride(1) != 1):
        raise RuntimeError(f"{name}: expected a 2-D tensor with unit column stride, got shape {tuple(t.shape)} strides {t.stride()}")
    return t


def padded_rows(m: int, n: int, dtype: torch.dtype, device, multiple: int = 64) -> torch.Tensor:
    """[m, n] view of a buffer whose row stride is rounded up to ``multiple`` elements: keeps rows 16-byte aligned for
    TMA / vector stores when n is odd-sized (the 6997-wide vocabulary logits -> stride 7040)."""
    ld = (n + multiple - 1) // multiple * multiple
    return torch.empty((m, ld), dtype=dtype, device=device)[:, :n]


def linear_fwd(x2d, w, bias=None, relu=False, out=None):
    """x2d [M,K] @ w[N,K]^T (+bias) -> [M,N] (nn.Linear semantics; w rows may be a slice of a packed weight)"""
    _rows2d(x2d, "linear_fwd.x")
    m, k = x2d.shape
    n = w.shape[0]
    y = torch.empty((m, n), dtype=x2d.dtype, device=x2d.device) if out is None else out
    gemm(x2d, w, y, m, n, k, trans_b=True, lda=x2d.stride(0), ldb=w.stride(0), ldc=y.stride(0), bias=bias, bias_mode=1, relu=relu)
    return y


def linear_dgrad(dy2d, w, out=None):
    """dx [M,K] = dy [M,N] @ w [N,K]"""
    _rows2d(dy2d, "linear_dgrad.dy")
    m, n = dy2d.shape
    k = w.shape[1]
    dx = torch.empty((m, k), dtype=dy2d.dtype, device=dy2d.device) if out is None else out
    gemm(dy2d, w, dx, m, k, n, lda=dy2d.stride(0), ldb=w.stride(0), ldc=dx.stride(0))
    return dx


def linear_wgrad(x2d, dy2d, dw, db=None, accumulate=True):
    """dw [N,K] (fp32, may be a row-slice view) += dy^T @ x ; db [N] += colsum(dy)"""
    _rows2d(x2d, "linear_wgrad.x"), _rows2d(dy2d, "linear_wgrad.dy")
    m, k = x2d.shape
    n = dy2d.shape[1]
    gemm(dy2d, x2d, dw, n, k, m, trans_a=True, lda=dy2d.stride(0), ldb=x2d.stride(0), ldc=dw.stride(0), accumulate=accumulate)
    if db is not None:
        colsum(dy2d, db, accumulate=accumulate)


def colsum(x2d, out, accumulate=True):
    rows, n = x2d.shape
    call("omr_colsum", dt_code(x2d.dtype), ptr(x2d), rows, n, x2d.stride(0), ptr(out), int(accumulate), stream_ptr())
    return out


class AttnSpec:
    """Mask description shared by attention forward and backward."""

    __slots__ = ("H", "hd", "scale", "causal", "window", "key_bias", "q_len", "kv_len", "quirk_mod", "dropout_p", "seed")

    def __init__(self, H, hd, causal=False, window=-1, key_bias=None, q_len=None, kv_len=None, quirk_mod=0, dropout_p=0.0,
                 seed=0):
        self.H, self.hd = H, hd
        # attention-probability dropout (train mode): the forward and the backward call must carry the same (p, seed)
        self.dropout_p, self.seed = float(dropout_p), int(seed)
        self.scale = 1.0 / math.sqrt(hd)
        self.causal, self.window = bool(causal), int(window) if window and window > 0 else 0
        self.key_bias, self.q_len, self.kv_len, self.quirk_mod = key_bias, q_len, kv_len, quirk_mod


def attn_spec_with_dropout(spec: "AttnSpec", p: float, seed: int) -> "AttnSpec":
    """copy of ``spec`` that applies attention-probability dropout (p, seed) in forward and backward"""
    return AttnSpec(spec.H, spec.hd, causal=spec.causal, window=spec.window, key_bias=spec.key_bias, q_len=spec.q_len,
                    kv_len=spec.kv_len, quirk_mod=spec.quirk_mod, dropout_p=p, seed=seed)


def _arm_attn_dropout(spec: "AttnSpec") -> None:
    if spec.dropout_p > 0.0:
        call("omr_attn_next_dropout", float(spec.dropout_p), int(spec.seed) & 0x7FFFFFFF, ptr(SEED_OFFSET_DEV))


def _view3(t, off, width):
    """(base pointer incl. column offset, batch stride, row stride) of columns [off, off+width) of a [B,T,W] buffer"""
    assert t.dim() == 3 and t.is_contiguous()
    return t.data_ptr() + off * t.element_size(), t.stride(0), t.stride(1)


def attn_fwd(qbuf, q_off, kbuf, k_off, vbuf, v_off, spec: AttnSpec):
    """q = qbuf[:, :, q_off:q_off+H*hd] etc. -> (o [B,Tq,H*hd], lse [B,H,Tq] fp32)"""
    b, tq, _ = qbuf.shape
    tk = kbuf.shape[1]
    d = spec.H * spec.hd
    o = torch.empty((b, tq, d), dtype=qbuf.dtype, device=qbuf.device)
    lse = torch.empty((b, spec.H, tq), dtype=torch.float32, device=qbuf.device)
    qp, qbs, qrs = _view3(qbuf, q_off, d)
    kp, kbs, krs = _view3(kbuf, k_off, d)
    vp, vbs, vrs = _view3(vbuf, v_off, d)
    _arm_attn_dropout(spec)
    call("omr_attn_fwd", dt_code(qbuf.dtype), qp, qbs, qrs, kp, kbs, krs, vp, vbs, vrs, ptr(o), o.stride(0), o.stride(1),
         ptr(lse), ptr(spec.key_bias), b, spec.H, tq, tk, spec.hd, spec.scale, int(spec.causal), spec.window,
         ptr(spec.q_len), ptr(spec.kv_len), spec.quirk_mod, stream_ptr())
    return o, lse


def attn_bwd(qbuf, q_off, kbuf, k_off, vbuf, v_off, o, do, lse, dqbuf, dq_off, dkbuf, dk_off, dvbuf, dv_off, spec: AttnSpec):
    """writes dq/dk/dv into column ranges of the given gradient buffers (same layouts as the inputs)"""
    _chk(do, "attn_bwd.do")
    b, tq, _ = qbuf.shape
    tk = kbuf.shape[1]
    d = spec.H * spec.hd
    qp, qbs, qrs = _view3(qbuf, q_off, d)
    kp, kbs, krs = _view3(kbuf, k_off, d)
    vp, vbs, vrs = _view3(vbuf, v_off, d)
    dqp, dqbs, dqrs = _view3(dqbuf, dq_off, d)
    dkp, dkbs, dkrs = _view3(dkbuf, dk_off, d)
    dvp, dvbs, dvrs = _view3(dvbuf, dv_off, d)
    delta = torch.empty(b * spec.H * tq * 65 + 4, dtype=torch.float32, device=qbuf.device)  # delta + fp32 dQ accumulators
    _arm_attn_dropout(spec)
    call("omr_attn_bwd", dt_code(qbuf.dtype), qp, qbs, qrs, kp, kbs, krs, vp, vbs, vrs, ptr(o), o.stride(0), o.stride(1),
         ptr(do), do.stride(0), do.stride(1), ptr(lse), dqp, dqbs, dqrs, dkp, dkbs, dkrs, dvp, dvbs, dvrs, ptr(delta),
         ptr(spec.key_bias), b, spec.H, tq, tk, spec.hd, spec.scale, int(spec.causal), spec.window, ptr(spec.q_len),
         ptr(spec.kv_len), spec.quirk_mod, stream_ptr())


def add_layernorm_fwd(x, res, gamma, beta, eps: float, save: bool):
    """y = LN(x + res); returns (y, s, stats) with s = x + res (compute dtype) when save else (y, None, None)"""
    _chk(x, "add_layernorm_fwd.x")
    d = x.shape[-1]
    rows = x.numel() // d
    y = torch.empty_like(x)
    s = torch.empty_like(x) if save else None
    stats = torch.empty((rows, 2), dtype=torch.float32, device=x.device) if save else None
    call("omr_add_layernorm_fwd", dt_code(x.dtype), ptr(x), ptr(res), ptr(gamma), ptr(beta), ptr(s), ptr(y), ptr(stats),
         rows, d, float(eps), stream_ptr())
    return y, s, stats


def layernorm_bwd(dy, s, stats, gamma, dgamma, dbeta):
    _chk(dy, "layernorm_bwd.dy")
    d = dy.shape[-1]
    rows = dy.numel() // d
    ds = torch.empty_like(dy)
    call("omr_layernorm_bwd", dt_code(dy.dtype), ptr(dy), ptr(s), ptr(stats), ptr(gamma), ptr(ds), ptr(dgamma), ptr(dbeta),
         rows, d, stream_ptr())
    return ds


def dropout_add_layernorm_fwd(x, res, gamma, beta, eps: float, save: bool, p: float, seed: int):
    """y = LN(dropout(x; p, seed) + res) in one kernel (x is NOT modified); same returns as add_layernorm_fwd"""
    _chk(x, "dropout_add_layernorm_fwd.x")
    d = x.shape[-1]
    rows = x.numel() // d
    y = torch.empty_like(x)
    s = torch.empty_like(x) if save else None
    stats = torch.empty((rows, 2), dtype=torch.float32, device=x.device) if save else None
    call("omr_dropout_add_layernorm_fwd", dt_code(x.dtype), ptr(x), ptr(res), ptr(gamma), ptr(beta), ptr(s), ptr(y), ptr(stats),
         rows, d, float(eps), float(p), int(seed), ptr(SEED_OFFSET_DEV), stream_ptr())
    return y, s, stats


def layernorm_bwd_dropout(dy, s, stats, gamma, dgamma, dbeta, p: float, seed: int, dbias=None):
    """-> (ds, da): ds as layernorm_bwd, da = dropout'(ds; p, seed) = the gradient of the dropped sublayer output;
    dbias (fp32 [D]) += column sums of da (the bias gradient of the linear layer in front of the dropout)"""
    _chk(dy, "layernorm_bwd_dropout.dy")
    d = dy.shape[-1]
    rows = dy.numel() // d
    ds = torch.empty_like(dy)
    da = torch.empty_like(dy)
    call("omr_layernorm_bwd_dropout", dt_code(dy.dtype), ptr(dy), ptr(s), ptr(stats), ptr(gamma), ptr(ds), ptr(da), ptr(dgamma),
         ptr(dbeta), rows, d, float(p), int(seed), ptr(SEED_OFFSET_DEV), ptr(dbias), stream_ptr())
    return ds, da


def mask_scale(dx, mask, scale: float):
    """dx *= (mask > 0 ? scale : 0) in place: backward of ReLU followed by dropout, mask = the dropped ReLU output"""
    _chk(dx, "mask_scale.dx"), _chk(mask, "mask_scale.mask")
    call("omr_mask_scale", dt_code(dx.dtype), ptr(dx), ptr(mask), float(scale), dx.numel(), stream_ptr())
    return dx


def ce_fwd(logits2d, targets, ignore_index: int):
    """logits [rows, V] (row stride may exceed V), targets int64 [rows] -> (loss_out fp32 [2] = (mean loss, n_valid), row_lse)"""
    rows, v = logits2d.shape
    dev = logits2d.device
    row_loss = torch.empty(rows, dtype=torch.float32, device=dev)
    row_lse = torch.empty(rows, dtype=torch.float32, device=dev)
    loss_out = torch.empty(2, dtype=torch.float32, device=dev)
    call("omr_ce_fwd", dt_code(logits2d.dtype), ptr(logits2d), logits2d.stride(0), ptr(targets), rows, v, ignore_index,
         ptr(row_loss), ptr(row_lse), stream_ptr())
    call("omr_ce_reduce", ptr(row_loss), ptr(targets), rows, ignore_index, ptr(loss_out), stream_ptr())
    return loss_out, row_lse


def ce_bwd(logits2d, targets, row_lse, loss_out, gscale, ignore_index: int, inplace=True):
    rows, v = logits2d.shape
    d = logits2d if inplace else torch.empty_like(logits2d)
    call("omr_ce_bwd", dt_code(logits2d.dtype), ptr(logits2d), logits2d.stride(0), ptr(targets), ptr(row_lse),
         ptr(loss_out), ptr(gscale), ptr(d), rows, v, ignore_index, stream_ptr())
    return d


# ---- classifier fused with the cross-entropy (csrc/projce_tc.cu): the [rows, V] logits are never written ---------------
def proj_ce_supported(dtype, d: int) -> bool:
    return bool(_lib.load().omr_proj_ce_supported(dt_code(dtype), int(d)))


def proj_ce_fwd(x2d, w, bias, targets, ignore_index: int):
    """x [rows, D], w [V, D], bias fp32 [V] | None, targets int64 [rows] -> (loss_out fp32 [2] = (mean loss, n_valid), row_lse)"""
    _chk(x2d, "proj_ce_fwd.x"), _chk(w, "proj_ce_fwd.w")
    rows, d = x2d.shape
    v = w.shape[0]
    dev = x2d.device
    row_loss = torch.empty(rows, dtype=torch.float32, device=dev)
    row_lse = torch.empty(rows, dtype=torch.float32, device=dev)
    loss_out = torch.empty(2, dtype=torch.float32, device=dev)
    call("omr_proj_ce_fwd", dt_code(x2d.dtype), ptr(x2d), x2d.stride(0), ptr(w), w.stride(0), ptr(bias), ptr(targets), rows, v, d,
         ignore_index, ptr(row_loss), ptr(row_lse), stream_ptr())
    call("omr_ce_reduce", ptr(row_loss), ptr(targets), rows, ignore_index, ptr(loss_out), stream_ptr())
    return loss_out, row_lse


def proj_ce_bwd_dx(x2d, w, bias, targets, row_lse, loss_out, gscale, ignore_index: int):
    """-> dL/dx [rows, D] (x's dtype)"""
    rows, d = x2d.shape
    dx = torch.empty((rows, d), dtype=x2d.dtype, device=x2d.device)
    call("omr_proj_ce_bwd_dx", dt_code(x2d.dtype), ptr(x2d), x2d.stride(0), ptr(w), w.stride(0), ptr(bias), ptr(targets), ptr(row_lse),
         ptr(loss_out), ptr(gscale), rows, w.shape[0], d, ignore_index, ptr(dx), dx.stride(0), stream_ptr())
    return dx


def proj_ce_bwd_dw(x2d, w, bias, targets, row_lse, loss_out, gscale, ignore_index: int, dw, db):
    """dw [V, D] fp32 and db [V] fp32 (or None) are accumulated into"""
    rows, d = x2d.shape
    call("omr_proj_ce_bwd_dw", dt_code(x2d.dtype), ptr(x2d), x2d.stride(0), ptr(w), w.stride(0), ptr(bias), ptr(targets), ptr(row_lse),
         ptr(loss_out), ptr(gscale), rows, w.shape[0], d, ignore_index, ptr(dw), ptr(db), stream_ptr())


# ---- greedy decode helpers ------------------------------------------------------------------------
def argmax_step(logits2d, tok, val, finished, eos_id, pad_id, out_tokens, out_vals, step, step_dev=None):
    b, v = logits2d.shape
    call("omr_argmax_step", dt_code(logits2d.dtype), ptr(logits2d), logits2d.stride(0), b, v, ptr(tok), ptr(val),
         ptr(finished), eos_id, pad_id, ptr(out_tokens), ptr(out_vals),
         out_tokens.shape[1] if out_tokens is not None else 0, step, ptr(step_dev), stream_ptr())


def kv_append(src_ptr, src_rs, cache, pos, dtype, pos_dev=None):
    b, tmax, width = cache.shape
    call("omr_kv_append", dt_code(dtype), src_ptr, src_rs, ptr(cache), b, tmax, width, pos, ptr(pos_dev), stream_ptr())


def attn_decode_ws_floats(b, h, hd=64):
    nsplit = max(1, -(-1184 // (b * h)))
    return b * h * nsplit * (hd + 2)


def attn_decode(q_ptr, q_bs, k_ptr, k_bs, k_rs, v_ptr, v_bs, v_rs, o, key_bias, ws, b, h, tk, hd, window, dtype,
                pos_dev=None):
    call("omr_attn_decode", dt_code(dtype), q_ptr, q_bs, k_ptr, k_bs, k_rs, v_ptr, v_bs, v_rs, ptr(o), o.stride(0),
         ptr(key_bias), key_bias.stride(0) if key_bias is not None else 0, ptr(ws), ws.numel(), b, h, tk, hd,
         1.0 / math.sqrt(hd), window if window and window > 0 else 0, ptr(pos_dev), stream_ptr())
    return o


def tick(counter_i32):
    """*counter += 1 on the stream (device-side step/position counter)"""
    call("omr_adam_tick", ptr(counter_i32), stream_ptr())


# ---- staging: collate, late fusion step, Levenshtein (csrc/staging.cu) ---------------------------------------
def pad_collate(flat, offsets, heights, widths, hmax, wmax, pad_value, height_reduction=16, width_reduction=8):
    """ragged fp32 samples (flat + offsets/heights/widths on the device) -> (x [B,1,Hmax,Wmax] fp32, n_frames int32 [B])"""
    _lib.require_cuda(flat, "pad_collate")
    b = heights.numel()
    x = torch.empty((b, 1, hmax, wmax), dtype=torch.float32, device=flat.device)
    nf = torch.empty((b,), dtype=torch.int32, device=flat.device)
    call("omr_pad_collate", ptr(flat), ptr(offsets), ptr(heights), ptr(widths), ptr(x), b, hmax, wmax, float(pad_value),
         ptr(nf), height_reduction, width_reduction, stream_ptr())
    return x, nf


def pad_transcripts(flat, offsets, t, pad_id=0):
    """ragged int64 transcripts -> (y_in, y_out) int64 [B,t] (transcript[:-1] / transcript[1:], zero padded)"""
    _lib.require_cuda(flat, "pad_transcripts")
    b = offsets.numel() - 1
    y_in = torch.empty((b, t), dtype=torch.int64, device=flat.device)
    y_out = torch.empty((b, t), dtype=torch.int64, device=flat.device)
    call("omr_pad_transcripts", ptr(flat), ptr(offsets), b, t, ptr(y_in), ptr(y_out), pad_id, stream_ptr())
    return y_in, y_out


def mix_argmax_step(logits_a, logits_b, alpha, tok, val, finished, eos_id, pad_id, out_tokens, out_vals, step, step_dev=None):
    b, v = logits_a.shape
    assert logits_b.shape == logits_a.shape and logits_b.dtype == logits_a.dtype
    call("omr_mix_argmax_step", dt_code(logits_a.dtype), ptr(logits_a), logits_a.stride(0), ptr(logits_b), logits_b.stride(0),
         b, v, float(alpha), ptr(tok), ptr(val), ptr(finished), eos_id, pad_id, ptr(out_tokens), ptr(out_vals),
         out_tokens.shape[1] if out_tokens is not None else 0, step, ptr(step_dev), stream_ptr())


def levenshtein(truth, truth_offsets, hyp, hyp_offsets, max_len):
    """-> (ed int32 [P], sums int64 [3] = {sum ed, sum truth length, #pairs with ed > 0}), all on the device"""
    _lib.require_cuda(truth_offsets, "levenshtein")
    p = truth_offsets.numel() - 1
    ed = torch.empty((p,), dtype=torch.int32, device=truth_offsets.device)
    sums = torch.zeros((3,), dtype=torch.int64, device=truth_offsets.device)
    call("omr_levenshtein", ptr(truth), ptr(truth_offsets), ptr(hyp), ptr(hyp_offsets), p, int(max_len), ptr(ed), ptr(sums),
         stream_ptr())
    return ed, sums
