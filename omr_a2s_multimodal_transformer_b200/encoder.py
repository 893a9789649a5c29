"""Convolutional image / spectrogram encoder on hand-written sm_100a kernels.

Drop-in for the reference ``src/transformer/encoder.py`` (same class names, constructor arguments,
state-dict keys and output semantics):

* ``Encoder(in_channels, dropout=0.5)``: 5 ``ConvBlock``s (1->16->32->64->128->128, conv3 strides
  (1,1),(2,2),(2,2),(2,2),(2,1)) + 4 ``DSCBlock``s with the conditional residual
  (reference encoder.py:241-291); output ``[B,256,ceil(H/16),ceil(W/8)]``.
* Activations are kept NHWC inside (the implicit-GEMM layout); the returned tensor has the reference's
  NCHW *shape* with channels-last strides, so ``pos_2d(x).flatten(2).permute(0,2,1).contiguous()``
  (reference model.py:145-147) is a zero-copy view chain.
* Backward is composed explicitly from the data-/weight-gradient kernels (one autograd node per
  encoder); parameter gradients are accumulated by the kernels straight into ``param.grad``.
"""
from __future__ import annotations

import os
import random
from typing import Callable, List, Optional, Tuple, Union

import torch
import torch.nn as nn

from . import ops
from .params import ConvParams, WeightCache, grad_buf, resolve_dtype

HEIGHT_REDUCTION = 16
WIDTH_REDUCTION = 8
IN_EPS = 1e-3  # nn.InstanceNorm2d(eps=0.001) at reference encoder.py:151-156, 210-215

Tape = Optional[List[Callable[[torch.Tensor], Optional[torch.Tensor]]]]


class _Ctx:
    """Per-call execution context: compute dtype, weight cache, backward tape, training flags."""

    # backward-time state shared by the closures of one encoder pass (set by _EncoderFn.backward): ``side`` = stream for
    # the work nothing later in the chain reads (weight / bias gradients), ``keep`` = tensors that work reads
    def __init__(self, dtype: torch.dtype, cache: WeightCache, tape: Tape, training: bool, dropout: "DropoutPlan",
                 bw: Optional[dict] = None):
        self.bw = bw if bw is not None else {"side": None, "keep": []}
        self.dtype, self.cache, self.tape, self.training, self.dropout = dtype, cache, tape, training, dropout


class DropoutPlan:
    """Train-time MixDropout decisions (reference encoder.py:87-104,160-179): ONE position per block
    (``random.randint(1, 3)``) and, there, element-wise Dropout(p) or channel-wise Dropout2d(p/2) with
    probability 1/2 each.  The Python RNG draws are made even in eval mode, like the reference."""

    def __init__(self, p: float, seed_fn: Callable[[], int]):
        self.p, self.seed_fn = p, seed_fn

    def draw(self) -> int:
        return random.randint(1, 3)

    def kind(self) -> bool:
        return random.random() < 0.5


FUSE_RELU_BWD = True  # tests flip this to compare against the unfused relu_bwd / dropout-backward kernels


def _off_chain(c: "_Ctx", fn, *tensors) -> None:
    """backward only: run fn() -- a weight / bias gradient, which nothing later in the chain reads -- on the side stream
    of this encoder pass, ordered after everything queued so far; ``tensors`` (its inputs) stay alive until the join in
    ``_EncoderFn.backward``.  Without a side stream fn() simply runs in place."""
    side = c.bw["side"]
    if side is None:
        fn()
        return
    c.bw["keep"].extend(tensors)
    side.wait_event(torch.cuda.current_stream(side.device).record_event())
    with torch.cuda.stream(side):
        fn()


class _Act:
    """Book-keeping for a ReLU output y (possibly followed by dropout) whose ONLY consumer can fuse the backward of that
    ReLU/dropout into its own gradient kernel (conv data-gradient epilogue, InstanceNorm backward): the consumer multiplies
    its dx by (x > 0 ? scale : 0), where x = the tensor it consumed -- zeros of x cover both inactive and dropped elements,
    scale is the dropout's 1/(1-p).  ``premasked`` is set by such a consumer at forward time; the producer's backward
    closures read it at backward time and skip their relu_bwd / dropout-backward kernels.  The masked dx IS the gradient of
    the producer's pre-activation, so the same consumer kernel also accumulates its column sums = the producer's BIAS
    gradient (``bias`` is registered by the producer, ``db_done`` set by the consumer at backward time)."""

    __slots__ = ("premasked", "scale", "bias", "db_done")

    def __init__(self) -> None:
        self.premasked = False
        self.scale = 1.0
        self.bias = None
        self.db_done = False


def _maybe_dropout(x: torch.Tensor, c: _Ctx, here: bool, act: Optional[_Act] = None, inplace_bwd: bool = True) -> torch.Tensor:
    """inplace_bwd=False: the incoming gradient has a second reader (the residual branch of a DSC block) and must not
    be overwritten by the dropout backward."""
    if not (here and c.training and c.dropout.p > 0.0):
        return x
    elementwise = c.dropout.kind()
    p = c.dropout.p if elementwise else c.dropout.p / 2
    seed = c.dropout.seed_fn()
    y = ops.dropout(x, p, seed, channelwise=not elementwise)
    if act is not None:
        act.scale = 1.0 / (1.0 - p)
    if c.tape is not None:

        def bwd(dy: torch.Tensor) -> torch.Tensor:
            if act is not None and act.premasked:
                return dy  # the consumer of y already applied (y > 0 ? 1/(1-p) : 0)
            return ops.dropout(dy, p, seed, channelwise=not elementwise, inplace=inplace_bwd)

        c.tape.append(bwd)
    return y


def _bias_target(act: Optional[_Act]) -> Optional[torch.Tensor]:
    """backward time: the bias-gradient buffer of the layer that produced the tensor described by ``act``, to be filled by
    the consumer's gradient kernel (and marked done), or None"""
    if act is None or act.bias is None or not act.bias.requires_grad:
        return None
    act.db_done = True
    return grad_buf(act.bias)


def _conv_step(x: torch.Tensor, cp: ConvParams, stride: Tuple[int, int], relu: bool, c: _Ctx, need_dx: bool,
               x_act: Optional[_Act] = None, y_act: Optional[_Act] = None, in_sums: Optional[torch.Tensor] = None,
               norm_link: Optional[dict] = None) -> torch.Tensor:
    """x_act: x is a ReLU(+dropout) output whose backward this conv's data gradient fuses; y_act: record for this conv's
    own ReLU output, filled in by whoever consumes it.  in_sums: receives the InstanceNorm statistics of the output (conv2
    of a block).  norm_link (conv3 of a block): x is an InstanceNorm output; the data gradient then also accumulates that
    norm's backward sums (``norm_link["x"]`` = the norm's input) into ``norm_link["bsums"]``."""
    wp = c.cache.get(cp.weight, "conv", c.dtype)
    y = ops.conv3x3_fwd(x, wp, cp.bias, stride, relu, in_sums=in_sums)
    if relu:
        ops._probe_relu(y)
    if y_act is not None:
        y_act.bias = cp.bias
    if c.tape is not None:
        in_hw = (x.shape[1], x.shape[2])
        fuse = FUSE_RELU_BWD and x_act is not None and need_dx
        if fuse:
            x_act.premasked = True

        def bwd(dy: torch.Tensor) -> Optional[torch.Tensor]:
            dz = dy
            if relu and not (y_act is not None and y_act.premasked):
                dz = ops.relu_bwd(y, dy, inplace=True)
            if cp.weight.requires_grad:
                gw = grad_buf(cp.weight)
                gb = None if (y_act is not None and y_act.db_done) else grad_buf(cp.bias)  # done by the consumer's kernel
                _off_chain(c, lambda: ops.conv3x3_wgrad(x, dz, gw, gb, stride, accumulate=True), x, dz)
            if not need_dx:
                return None
            wt = c.cache.get(cp.weight, "convT", c.dtype)
            if fuse:
                return ops.conv3x3_dgrad(dz, wt, in_hw, stride, mask=x, mask_scale=x_act.scale, colsum=_bias_target(x_act))
            if norm_link is not None and FUSE_NORM_SUMS:
                bs = ops.in_sums_buffer(x.shape[0], x.shape[3], x.device)
                norm_link["bsums"] = bs
                return ops.conv3x3_dgrad(dz, wt, in_hw, stride, in_x=norm_link["x"], in_bsums=bs)
            return ops.conv3x3_dgrad(dz, wt, in_hw, stride)

        c.tape.append(bwd)
    return y


FUSE_NORM_SUMS = True  # tests flip this to compare against the separate InstanceNorm statistics / reduction passes


def _instnorm_step(x: torch.Tensor, c: _Ctx, x_act: Optional[_Act] = None, sums: Optional[torch.Tensor] = None,
                   norm_link: Optional[dict] = None) -> torch.Tensor:
    """sums: (sum x, sum x^2) already accumulated by the convolution that produced x; norm_link: filled by the consumer
    convolution's data gradient with the backward sums (see _conv_step)"""
    y, stats = ops.instnorm_fwd(x, IN_EPS, sums=sums)
    if norm_link is not None:
        norm_link["x"] = x
    if c.tape is not None:

        def bwd(dy: torch.Tensor) -> torch.Tensor:
            bs = norm_link.pop("bsums", None) if norm_link is not None else None
            if FUSE_RELU_BWD and x_act is not None:
                return ops.instnorm_bwd(dy, x, stats, relu_mask=True, mask_scale=x_act.scale, sums=bs, colsum=_bias_target(x_act))
            return ops.instnorm_bwd(dy, x, stats, sums=bs)

        if FUSE_RELU_BWD and x_act is not None:
            x_act.premasked = True
        c.tape.append(bwd)
    return y


def _dw_step(x: torch.Tensor, cp: ConvParams, c: _Ctx) -> torch.Tensor:
    wp = c.cache.get(cp.weight, "dw", c.dtype)
    y = ops.dwconv3x3_fwd(x, wp, cp.bias)
    if c.tape is not None:

        def bwd(dy: torch.Tensor) -> torch.Tensor:
            if cp.weight.requires_grad:
                gw, gb = grad_buf(cp.weight), grad_buf(cp.bias)
                _off_chain(c, lambda: ops.dwconv3x3_wgrad(x, dy, gw, gb, accumulate=True), x, dy)
            return ops.dwconv3x3_dgrad(dy, wp)

        c.tape.append(bwd)
    return y


def _pw_step(x: torch.Tensor, cp: ConvParams, relu: bool, c: _Ctx) -> torch.Tensor:
    n, h, w, ci = x.shape
    co = cp.out_channels
    wm = c.cache.get(cp.weight, "mat", c.dtype)
    x2 = x.view(-1, ci)
    y2 = ops.linear_fwd(x2, wm, cp.bias, relu=relu)
    if relu:
        ops._probe_relu(y2.view(n, h, w, co))
    if c.tape is not None:

        def bwd(dy: torch.Tensor) -> torch.Tensor:
            dz = dy.view(-1, co)
            if relu:
                dz = ops.relu_bwd(y2, dz, inplace=True)
            if cp.weight.requires_grad:
                gw, gb = grad_buf(cp.weight).view(co, ci), grad_buf(cp.bias)
                _off_chain(c, lambda: ops.linear_wgrad(x2, dz, gw, gb, accumulate=True), x2, dz)
            return ops.linear_dgrad(dz, wm).view(n, h, w, ci)

        c.tape.append(bwd)
    return y2.view(n, h, w, co)


class DepthSepConv2D(nn.Module):
    """Depthwise 3x3 (pad 1, stride 1) + pointwise 1x1 (reference encoder.py:12-84; the only
    configuration the reference instantiates is kernel (3,3), stride (1,1), no inner activation)."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size: Tuple[int, int] = (3, 3), activation=None,
                 padding: Union[bool, Tuple[int, int]] = True, stride: Union[int, Tuple[int, int]] = (1, 1),
                 dilation: Union[int, Tuple[int, int]] = (1, 1)):
        super().__init__()
        if tuple(kernel_size) != (3, 3) or tuple(_pair(stride)) != (1, 1) or tuple(_pair(dilation)) != (1, 1) or activation:
            raise NotImplementedError("DepthSepConv2D kernels cover the reference's configuration: 3x3, stride 1, dilation 1")
        self.depth_conv = ConvParams(in_channels, in_channels, (3, 3), groups=in_channels)
        self.point_conv = ConvParams(in_channels, out_channels, (1, 1))

    def _run(self, x: torch.Tensor, relu: bool, c: _Ctx) -> torch.Tensor:
        return _pw_step(_dw_step(x, self.depth_conv, c), self.point_conv, relu, c)


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


class MixDropout(nn.Module):
    """Kept for surface compatibility (reference encoder.py:87-104); the decisions are drawn by
    ``DropoutPlan`` and applied by the fused dropout kernel."""

    def __init__(self, dropout_prob: float = 0.4, dropout_2d_prob: float = 0.2):
        super().__init__()
        self.dropout_prob, self.dropout_2d_prob = dropout_prob, dropout_2d_prob


class ConvBlock(nn.Module):
    """conv3x3+ReLU, conv3x3+ReLU, InstanceNorm, strided conv3x3+ReLU (reference encoder.py:107-181)."""

    def __init__(self, in_c: int, out_c: int, stride=(1, 1), kernel: int = 3, activation=None, dropout: float = 0.5):
        super().__init__()
        if kernel != 3:
            raise NotImplementedError("ConvBlock kernels cover the reference's configuration: kernel 3")
        self.stride = _pair(stride)
        self.conv1 = ConvParams(in_c, out_c, (3, 3))
        self.conv2 = ConvParams(out_c, out_c, (3, 3))
        self.conv3 = ConvParams(out_c, out_c, (3, 3))
        self.dropout = MixDropout(dropout_prob=dropout, dropout_2d_prob=dropout / 2)

    def _run(self, x: torch.Tensor, c: _Ctx, need_dx: bool, x_act: Optional[_Act] = None,
             out_act: Optional[_Act] = None) -> torch.Tensor:
        """x_act: record of the ReLU output x comes from (previous block), out_act: record for this block's output, to
        be completed by the next block's first convolution."""
        pos = c.dropout.draw()
        a1, a2 = _Act(), _Act()
        x = _conv_step(x, self.conv1, (1, 1), True, c, need_dx, x_act=x_act, y_act=a1)
        x = _maybe_dropout(x, c, pos == 1, a1)
        # the InstanceNorm statistics come out of conv2's epilogue unless a dropout sits between the two (the norm then
        # sees the dropped tensor); the norm's backward sums come out of conv3's data gradient
        drop2 = pos == 2 and c.training and c.dropout.p > 0.0
        sums = ops.in_sums_buffer(x.shape[0], self.conv2.out_channels, x.device) if (FUSE_NORM_SUMS and not drop2) else None
        link: dict = {}
        x = _conv_step(x, self.conv2, (1, 1), True, c, True, x_act=a1, y_act=a2, in_sums=sums)
        x = _maybe_dropout(x, c, pos == 2, a2)
        x = _instnorm_step(x, c, x_act=a2, sums=sums, norm_link=link)
        x = _conv_step(x, self.conv3, self.stride, True, c, True, y_act=out_act, norm_link=link)
        x = _maybe_dropout(x, c, pos == 3, out_act)
        return x


class DSCBlock(nn.Module):
    """3x DepthSepConv2D, ReLU after the first two, InstanceNorm before the third
    (reference encoder.py:184-238)."""

    def __init__(self, in_c: int, out_c: int, stride=(2, 1), activation=None, dropout: float = 0.5):
        super().__init__()
        if _pair(stride) != (1, 1):
            raise NotImplementedError("DSCBlock kernels cover the reference's configuration: stride (1,1)")
        self.conv1 = DepthSepConv2D(in_c, out_c)
        self.conv2 = DepthSepConv2D(out_c, out_c)
        self.conv3 = DepthSepConv2D(out_c, out_c)
        self.dropout = MixDropout(dropout_prob=dropout, dropout_2d_prob=dropout / 2)

    def _run(self, x: torch.Tensor, c: _Ctx) -> torch.Tensor:
        pos = c.dropout.draw()
        x = self.conv1._run(x, True, c)
        x = _maybe_dropout(x, c, pos == 1)
        x = self.conv2._run(x, True, c)
        x = _maybe_dropout(x, c, pos == 2)
        x = _instnorm_step(x, c)
        x = self.conv3._run(x, False, c)
        # the block's output gradient is also the gradient of the residual branch (Encoder._run): keep it intact
        x = _maybe_dropout(x, c, pos == 3, inplace_bwd=False)
        return x


class _EncoderFn(torch.autograd.Function):
    """One autograd node for the whole encoder: forward records a tape of kernel-level backward
    steps, backward replays it.  Parameter gradients are written by the kernels into ``param.grad``
    (the parameters are listed as inputs only so that the node is part of the graph)."""

    @staticmethod
    def forward(ctx, x_nhwc: torch.Tensor, enc: "Encoder", dtype: torch.dtype, training: bool, *params):
        tape: List = []
        bw = {"side": None, "keep": []}
        y = enc._run(x_nhwc, dtype, tape, training, bw)
        ctx.tape = tape
        ctx.bw = bw
        ctx.enc = enc
        return y

    @staticmethod
    def backward(ctx, dy: torch.Tensor):
        tape, ctx.tape = ctx.tape, None
        if tape is None:
            raise RuntimeError("encoder backward called twice (activations are released after the first pass)")
        g: Optional[torch.Tensor] = dy.contiguous().clone()
        # the data-gradient chain stays on this stream; the weight gradients (about as much kernel time again) go to a
        # side stream and fill the SMs the chain leaves idle (OMR_OVERLAP_ENCODER_WGRAD=0: one stream)
        side = ctx.enc._wgrad_stream(dy.device)
        cur = torch.cuda.current_stream(dy.device) if side is not None else None
        ctx.bw["side"] = side
        if side is not None:
            side.wait_stream(cur)
        while tape:
            step = tape.pop()
            g = step(g)
        if side is not None:
            cur.wait_stream(side)
        ctx.bw["side"] = None
        ctx.bw["keep"].clear()
        cb = getattr(ctx.enc, "_bwd_done_cb", None)
        if cb is not None:  # data-parallel: this encoder's gradient bucket is complete (ddp.py)
            cb()
        return (None, None, None, None) + tuple(None for _ in ctx.needs_input_grad[4:])


class Encoder(nn.Module):
    """Reference ``Encoder(in_channels, dropout=0.5)`` (encoder.py:241-291) on sm_100a kernels."""

    def __init__(self, in_channels: int, dropout: float = 0.5, out_channels: int = 256):
        super().__init__()
        self.in_channels = in_channels
        self.dropout_p = dropout
        self.conv_blocks = nn.ModuleList(
            [
                ConvBlock(in_channels, 16, stride=(1, 1), dropout=dropout),
                ConvBlock(16, 32, stride=(2, 2), dropout=dropout),
                ConvBlock(32, 64, stride=(2, 2), dropout=dropout),
                ConvBlock(64, 128, stride=(2, 2), dropout=dropout),
                ConvBlock(128, 128, stride=(2, 1), dropout=dropout),
            ]
        )
        self.dscblocks = nn.ModuleList(
            [
                DSCBlock(128, 128, stride=(1, 1), dropout=dropout),
                DSCBlock(128, 128, stride=(1, 1), dropout=dropout),
                DSCBlock(128, 128, stride=(1, 1), dropout=dropout),
                DSCBlock(128, out_channels, stride=(1, 1), dropout=dropout),
            ]
        )
        self.compute_dtype: Optional[torch.dtype] = None
        self._wcache = WeightCache()
        self._seed_state = 0x1234567

    # -- plumbing ---------------------------------------------------------------------------------
    def _next_seed(self) -> int:
        self._seed_state = (self._seed_state * 6364136223846793005 + 1442695040888963407) & 0xFFFFFFFFFFFFFFFF
        return (self._seed_state >> 17) & 0x7FFFFFFF

    def _wgrad_stream(self, device) -> Optional["torch.cuda.Stream"]:
        if os.environ.get("OMR_OVERLAP_ENCODER_WGRAD", "1") == "0" or torch.device(device).type != "cuda":
            return None
        side = getattr(self, "_wg_side", None)
        if side is None or side.device != torch.device(device):
            side = torch.cuda.Stream(device=device)
            self._wg_side = side
        return side

    def _run(self, x: torch.Tensor, dtype: torch.dtype, tape: Tape, training: bool, bw: Optional[dict] = None) -> torch.Tensor:
        c = _Ctx(dtype, self._wcache, tape, training, DropoutPlan(self.dropout_p, self._next_seed), bw)
        prev_act: Optional[_Act] = None
        nblk = len(self.conv_blocks)
        for i, blk in enumerate(self.conv_blocks):
            # the output of the last block feeds the DSC stack (depthwise conv + residual add: several consumers), so
            # only blocks 0..n-2 hand their output record to the next block's first convolution
            out_act = _Act() if i + 1 < nblk else None
            x = blk._run(x, c, need_dx=(i > 0), x_act=prev_act, out_act=out_act)
            prev_act = out_act
        for blk in self.dscblocks:
            if tape is None:
                xt = blk._run(x, c)
                x = ops.add(x, xt) if x.shape == xt.shape else xt
            else:
                inner: List = []
                ci = _Ctx(dtype, self._wcache, inner, training, c.dropout, c.bw)
                xt = blk._run(x, ci)
                residual = x.shape == xt.shape

                def bwd(dy: torch.Tensor, inner=inner, residual=residual) -> torch.Tensor:
                    g = dy
                    while inner:
                        g = inner.pop()(g)
                    return ops.add(g, dy) if residual else g

                tape.append(bwd)
                x = ops.add(x, xt) if residual else xt
        return x

    def forward_nhwc(self, x: torch.Tensor) -> torch.Tensor:
        """x [B,C_in,H,W] float32 -> features [B,h,w,C_out] (NHWC, compute dtype)."""
        ops._lib.require_cuda(x, "Encoder.forward")
        dtype = resolve_dtype(self.compute_dtype)
        b, cin, h, w = x.shape
        if cin != self.in_channels:
            raise RuntimeError(f"Encoder expects {self.in_channels} input channel(s), got {cin}")
        x_nhwc = x.reshape(b, h, w, 1) if cin == 1 else x.permute(0, 2, 3, 1)
        x_nhwc = ops.cast(x_nhwc.contiguous().float(), dtype)
        params = [p for p in self.parameters()]
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            return _EncoderFn.apply(x_nhwc, self, dtype, self.training, *params)
        return self._run(x_nhwc, dtype, None, self.training)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """Reference semantics: [B,C_in,H,W] -> [B,256,ceil(H/16),ceil(W/8)] (channels-last strides)."""
        return self.forward_nhwc(x).permute(0, 3, 1, 2)
