"""Transformer decoder (embedding + 1-D PE + 8 post-norm layers + vocabulary classifier) on
hand-written sm_100a kernels.

Drop-in for the reference ``src/transformer/decoder.py``: same ``Decoder`` constructor arguments,
``forward(tgt, memory, memory_len) -> logits [B,V,T]``, same state-dict keys as the
``nn.Embedding`` / ``nn.TransformerDecoder`` / ``nn.Conv1d`` it replaces, and the same mask algebra
(SURVEY.md appendix B): bool memory masks exclude keys, integer ``memory_len`` becomes the
reference's *additive* +1.0 float mask, ``tgt == 0`` adds +1.0 to padded target keys whenever a
memory mask is present, causal or sliding-window structure on the self-attention.

One autograd node covers the whole layer stack; its backward is composed from the gradient kernels
and accumulates parameter gradients straight into ``param.grad``.
"""
from __future__ import annotations

import os
from typing import List, Optional, Tuple

import torch
import torch.nn as nn

from . import ops
from .ops import AttnSpec
from .params import ConvParams, LayerNormParams, LinearParams, MHAParams, WeightCache, grad_buf, resolve_dtype


class PositionalEncoding1D(nn.Module):
    """Sinusoidal PE buffer ``pe [1,max_len,D]`` (reference decoder.py:7-32)."""

    def __init__(self, max_len: int, emb_dim: int, dropout_p: float = 0.1):
        super().__init__()
        self.dropout_p = dropout_p
        pos = torch.arange(max_len).unsqueeze(1)
        den = torch.pow(10000, torch.arange(0, emb_dim, 2) / emb_dim)
        pe = torch.zeros(1, max_len, emb_dim)
        pe[0, :, 0::2] = torch.sin(pos / den)
        pe[0, :, 1::2] = torch.cos(pos / den)
        self.register_buffer("pe", pe)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x [B,T,D] + pe[:, :T] (eval semantics; inside ``Decoder`` this add is fused with the gather)."""
        b, t, d = x.shape
        x = x.contiguous()
        return ops.pe2d_add(x.view(b, 1, t, d), self.pe.view(1, -1, d), torch.empty_like(x), 0)


class _Embedding(nn.Module):
    def __init__(self, num_embeddings: int, embedding_dim: int, padding_idx: int):
        super().__init__()
        self.num_embeddings, self.embedding_dim, self.padding_idx = num_embeddings, embedding_dim, padding_idx
        self.weight = nn.Parameter(torch.empty(num_embeddings, embedding_dim))
        nn.init.normal_(self.weight)
        with torch.no_grad():
            self.weight[padding_idx].zero_()


class _DecoderLayer(nn.Module):
    """Parameters of one nn.TransformerDecoderLayer (post-norm, ReLU FFN)."""

    def __init__(self, d_model: int, nhead: int, dim_feedforward: int):
        super().__init__()
        self.self_attn = MHAParams(d_model, nhead)
        self.multihead_attn = MHAParams(d_model, nhead)
        self.linear1 = LinearParams(d_model, dim_feedforward)
        self.linear2 = LinearParams(dim_feedforward, d_model)
        self.norm1 = LayerNormParams(d_model)
        self.norm2 = LayerNormParams(d_model)
        self.norm3 = LayerNormParams(d_model)


class _TransformerDecoder(nn.Module):
    def __init__(self, d_model: int, nhead: int, dim_feedforward: int, num_layers: int):
        super().__init__()
        first = _DecoderLayer(d_model, nhead, dim_feedforward)
        layers = [first]
        for _ in range(num_layers - 1):  # nn.TransformerDecoder deep-copies one layer: identical init
            nxt = _DecoderLayer(d_model, nhead, dim_feedforward)
            nxt.load_state_dict(first.state_dict())
            layers.append(nxt)
        self.layers = nn.ModuleList(layers)
        self.num_layers = num_layers


class _DecoderStackFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, memory: torch.Tensor, dec: "Decoder", tgt: torch.Tensor, mem_bias, tgt_bias, dtype, training: bool,
                *params):
        tape: List = []
        y = dec._run_stack(tgt, memory, mem_bias, tgt_bias, dtype, tape, training)
        ctx.tape = tape
        ctx.dec = dec
        ctx.mem_shape = memory.shape
        return y

    @staticmethod
    def backward(ctx, dy: torch.Tensor):
        tape, ctx.tape = ctx.tape, None
        if tape is None:
            raise RuntimeError("decoder backward called twice (activations are released after the first pass)")
        need_dmem = ctx.needs_input_grad[0]
        side = ctx.dec._side_stream(dy.device)
        # "keep": tensors read by work queued on the side stream; they must outlive it (the allocator only orders reuse
        # against the stream a block was allocated on), so they are released after the join below
        state = {"g": dy.contiguous().clone(), "dmem": None, "need_dmem": need_dmem, "side": side, "keep": []}
        if side is not None:
            side.wait_stream(torch.cuda.current_stream(dy.device))
        while tape:
            tape.pop()(state)
        if side is not None:
            torch.cuda.current_stream(dy.device).wait_stream(side)
        state["keep"].clear()
        ctx.dec._side_keep = []
        cb = getattr(ctx.dec, "_bwd_done_cb", None)
        if cb is not None:  # data-parallel: the decoder's gradient bucket is complete (ddp.py)
            cb()
        return (state["dmem"], None, None, None, None, None, None) + tuple(None for _ in ctx.needs_input_grad[7:])


class _LogitsFn(torch.autograd.Function):
    """hidden [B,T,D] -> logits [B,V,T] (the reference's permute + Conv1d(k=1), decoder.py:145-146)."""

    @staticmethod
    def forward(ctx, hidden: torch.Tensor, dec: "Decoder", dtype, weight, bias):
        ctx.dec, ctx.dtype = dec, dtype
        ctx.save_for_backward(hidden)
        return dec._logits_bvt(hidden, dtype)

    @staticmethod
    def backward(ctx, dlogits: torch.Tensor):
        (hidden,) = ctx.saved_tensors
        dec, dtype = ctx.dec, ctx.dtype
        dl = dlogits.contiguous()
        b, v, t = dl.shape
        d = hidden.shape[2]
        wm = dec._wcache.get(dec.out_layer.weight, "mat", dtype)
        dh = torch.empty_like(hidden)
        ops.gemm(dl, wm, dh, t, d, v, trans_a=True, lda=t, ldb=d, ldc=d, batch=b, stride_a=v * t, stride_c=t * d)
        if dec.out_layer.weight.requires_grad:
            dw = grad_buf(dec.out_layer.weight).view(v, d)
            db = grad_buf(dec.out_layer.bias)
            ones = torch.ones(t, dtype=dtype, device=dl.device)
            for i in range(b):
                ops.gemm(dl[i], hidden[i], dw, v, d, t, lda=t, ldb=d, ldc=d, accumulate=True)
                ops.gemm(dl[i], ones, db, v, 1, t, lda=t, ldb=1, ldc=1, accumulate=True)
        return dh, None, None, None, None


class _ProjCEFn(torch.autograd.Function):
    """hidden [B,T,D], targets [B,T] -> mean softmax cross-entropy over non-ignored targets
    (classifier + CrossEntropyLoss(ignore_index) of reference model.py:109,166,444,588).

    bf16, d_model 256 (``ops.proj_ce_supported``): the classifier GEMM and the cross-entropy are ONE tensor-core kernel per
    direction (csrc/projce_tc.cu) -- the [B*T, V] logits are never written; forward keeps one lse per row, the backward
    recomputes the score tiles in TMEM.  ``OMR_FUSE_PROJ_CE=0`` (or fp32 / another width) takes the round-1 path: GEMM ->
    logits -> row-wise lse -> in-place dlogits -> two GEMMs."""

    @staticmethod
    def forward(ctx, hidden: torch.Tensor, dec: "Decoder", targets: torch.Tensor, ignore_index: int, dtype, weight, bias):
        b, t, d = hidden.shape
        wm = dec._wcache.get(dec.out_layer.weight, "mat", dtype)
        h2 = hidden.reshape(b * t, d)
        tg = targets.reshape(-1).contiguous()
        ctx.dec, ctx.dtype, ctx.ignore = dec, dtype, ignore_index
        ctx.shape = (b, t, d)
        ctx.fused = os.environ.get("OMR_FUSE_PROJ_CE", "1") != "0" and ops.proj_ce_supported(dtype, d)
        if ctx.fused:
            loss_out, row_lse = ops.proj_ce_fwd(h2, wm, dec.out_layer.bias, tg, ignore_index)
            ctx.save_for_backward(h2, tg, row_lse, loss_out)
            return loss_out[0].clone()
        # logits rows padded to a multiple of 64 elements (6997 -> 7040): 16-byte aligned rows for TMA / vector stores
        logits = ops.linear_fwd(h2, wm, dec.out_layer.bias, out=ops.padded_rows(b * t, wm.shape[0], dtype, h2.device))
        loss_out, row_lse = ops.ce_fwd(logits, tg, ignore_index)
        ctx.save_for_backward(h2, logits, tg, row_lse, loss_out)
        return loss_out[0].clone()

    @staticmethod
    def backward(ctx, gloss: torch.Tensor):
        dec, dtype = ctx.dec, ctx.dtype
        g = gloss.reshape(1).float().contiguous()
        wm = dec._wcache.get(dec.out_layer.weight, "mat", dtype)
        v, d = wm.shape
        train_w = dec.out_layer.weight.requires_grad
        # the classifier's weight gradient is off the chain; the decoder stack's backward, which follows whenever the hidden
        # state needs a gradient, joins the side stream and releases the tensors kept here
        side = dec._side_stream(gloss.device) if (train_w and ctx.needs_input_grad[0]) else None
        if ctx.fused:
            h2, tg, row_lse, loss_out = ctx.saved_tensors
            bias = dec.out_layer.bias

            def wgrad():
                ops.proj_ce_bwd_dw(h2, wm, bias, tg, row_lse, loss_out, g, ctx.ignore, grad_buf(dec.out_layer.weight).view(v, d),
                                   grad_buf(bias))

            if train_w and side is not None:
                dec._side_keep = [h2, tg, row_lse, loss_out, g, wm]
                side.wait_event(torch.cuda.current_stream(gloss.device).record_event())
                with torch.cuda.stream(side):
                    wgrad()
            dh = ops.proj_ce_bwd_dx(h2, wm, bias, tg, row_lse, loss_out, g, ctx.ignore)
            if train_w and side is None:
                wgrad()
            return dh.view(ctx.shape), None, None, None, None, None, None
        h2, logits, tg, row_lse, loss_out = ctx.saved_tensors
        dl = ops.ce_bwd(logits, tg, row_lse, loss_out, g, ctx.ignore, inplace=True)
        dh = ops.linear_dgrad(dl, wm)
        if train_w:
            gw, gb = grad_buf(dec.out_layer.weight).view(v, d), grad_buf(dec.out_layer.bias)
            if side is None:
                ops.linear_wgrad(h2, dl, gw, gb)
            else:
                dec._side_keep = [h2, dl]
                side.wait_event(torch.cuda.current_stream(dl.device).record_event())
                with torch.cuda.stream(side):
                    ops.linear_wgrad(h2, dl, gw, gb)
        return dh.view(ctx.shape), None, None, None, None, None, None


class Decoder(nn.Module):
    """Reference ``Decoder`` (decoder.py:35-148) on sm_100a kernels."""

    def __init__(self, output_size: int, max_seq_len: int, num_embeddings: int, embedding_dim: int = 256,
                 padding_idx: int = 0, ff_dim: int = 256, dropout_p: float = 0.1, nhead: int = 4,
                 num_transformer_layers: int = 8, attn_window: int = -1):
        super().__init__()
        if embedding_dim % nhead != 0 or embedding_dim // nhead != 64:
            raise NotImplementedError("attention kernels are specialised for head_dim 64 (the reference's 256/4)")
        self.embedding = _Embedding(num_embeddings, embedding_dim, padding_idx)
        self.pos_1d = PositionalEncoding1D(max_seq_len, embedding_dim, dropout_p)
        self.attn_window = attn_window
        self.transformer_decoder = _TransformerDecoder(embedding_dim, nhead, ff_dim, num_transformer_layers)
        self.out_layer = ConvParams(embedding_dim, output_size, (1,))
        self.d_model, self.nhead, self.ff_dim, self.dropout_p = embedding_dim, nhead, ff_dim, dropout_p
        self.output_size, self.max_seq_len, self.padding_idx = output_size, max_seq_len, padding_idx
        self.compute_dtype: Optional[torch.dtype] = None
        self._wcache = WeightCache()
        self._seed_state = 0x7654321

    # ---- reference-compatible mask builders (API surface; the kernels take key-bias vectors) -----
    def get_memory_key_padding_mask(self, memory: torch.Tensor, memory_len: Optional[torch.Tensor] = None):
        """decoder.py:150-189: None -> None; bool [B,S] -> clone; lengths -> FLOAT 0/1 mask."""
        if memory_len is None:
            return None
        if memory_len.dtype == torch.bool:
            assert memory_len.shape[0] == memory.shape[0], (
                f"Different batch sizes for memory and memory_len: {memory.shape[0]} != {memory_len.shape[0]}"
            )
            assert memory_len.shape[1] == memory.shape[1], (
                f"Different sequence lengths for memory and memory_len: {memory.shape[1]} != {memory_len.shape[1]}"
            )
            return memory_len.clone()
        pos = torch.arange(memory.shape[1], device=memory.device).unsqueeze(0)
        return (pos >= memory_len.to(memory.device).unsqueeze(1)).to(torch.float32)

    @staticmethod
    def create_variable_window_mask(size: int, window_size: int, dtype=torch.float32, device=torch.device("cpu")):
        """decoder.py:191-217."""
        i = torch.arange(size, device=device).unsqueeze(1)
        j = torch.arange(size, device=device).unsqueeze(0)
        ok = j <= i
        if window_size < size:
            ok = ok & (j >= i - window_size)
        return torch.zeros(size, size, dtype=dtype, device=device).masked_fill(~ok, float("-inf"))

    def get_tgt_masks(self, tgt: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """decoder.py:219-254."""
        t = tgt.shape[1]
        w = self.attn_window if self.attn_window > 0 else t
        return self.create_variable_window_mask(t, w, device=tgt.device), (tgt == 0).to(torch.float32)

    # ---- kernels ---------------------------------------------------------------------------------------
    def _side_stream(self, device) -> Optional["torch.cuda.Stream"]:
        """The decoder is one dependent chain of small kernels (most of them less than one wave of CTAs); the work that
        is NOT on that chain -- the 8 cross-K/V projections of the memory in the forward, every weight/bias gradient and
        the memory gradient in the backward -- is issued on this side stream and fills the idle SMs.
        ``OMR_OVERLAP_DECODER=0`` keeps everything on one stream."""
        if os.environ.get("OMR_OVERLAP_DECODER", "1") == "0" or torch.device(device).type != "cuda":
            return None
        side = getattr(self, "_side", None)
        if side is None or side.device != torch.device(device):
            side = torch.cuda.Stream(device=device)
            self._side = side
        return side

    def _next_seed(self) -> int:
        self._seed_state = (self._seed_state * 6364136223846793005 + 1442695040888963407) & 0xFFFFFFFFFFFFFFFF
        return (self._seed_state >> 17) & 0x7FFFFFFF

    def _key_biases(self, tgt: torch.Tensor, memory: torch.Tensor, memory_len):
        """(memory key bias [B,S] fp32 | None, target key bias [B,T] fp32 | None) -- decoder.py:128-132."""
        if memory_len is None:
            return None, None
        b, s = memory.shape[0], memory.shape[1]
        if memory_len.dtype == torch.bool:
            assert memory_len.shape[0] == b, f"Different batch sizes for memory and memory_len: {b} != {memory_len.shape[0]}"
            assert memory_len.shape[1] == s, (
                f"Different sequence lengths for memory and memory_len: {s} != {memory_len.shape[1]}"
            )
            mem_bias = torch.zeros((b, s), dtype=torch.float32, device=memory.device).masked_fill_(
                memory_len.to(memory.device), float("-inf"))
        elif memory_len.dtype == torch.float32 and memory_len.dim() == 2:
            mem_bias = memory_len.to(memory.device).contiguous()  # already a key bias (internal fast path)
        else:
            lens = memory_len.to(device=memory.device, dtype=torch.int32).contiguous()
            mem_bias = torch.empty((b, s), dtype=torch.float32, device=memory.device)
            ops.key_bias_from_lengths(mem_bias, lens, 0, s, 1.0)
        tgt_bias = ops.key_bias_from_tokens(tgt.contiguous(), 0, 1.0)  # literal 0, decoder.py:253
        return mem_bias, tgt_bias

    def _drop(self, x: torch.Tensor, tape, training: bool) -> torch.Tensor:
        """nn.Dropout(p=dropout_p) of the embedding / residual branches (train mode only)."""
        if not training or self.dropout_p <= 0.0:
            return x
        seed = self._next_seed()
        p = self.dropout_p
        y = ops.dropout(x, p, seed, inplace=True)
        if tape is not None:
            tape.append(lambda st: st.__setitem__("g", ops.dropout(st["g"], p, seed, inplace=True)))
        return y

    def _run_stack(self, tgt, memory, mem_bias, tgt_bias, dtype, tape, training: bool) -> torch.Tensor:
        """embedding+PE and the layer stack; returns hidden [B,T,D].  ``tape`` (or None) receives the
        backward steps, each a callable on the state dict {g: dL/dx, dmem, need_dmem}."""
        c = self._wcache
        b, t = tgt.shape
        s = memory.shape[1]
        d, h = self.d_model, self.nhead
        hd = d // h
        mem2 = memory.reshape(b * s, d)
        table = c.get(self.embedding.weight, "mat", dtype)
        tgt = tgt.contiguous()
        x = ops.embed_pe_fwd(tgt, table, self.pos_1d.pe.view(-1, d))
        if tape is not None and self.embedding.weight.requires_grad:
            tape.append(lambda st: ops.embed_bwd(tgt, st["g"], grad_buf(self.embedding.weight), self.padding_idx))
        x = self._drop(x, tape, training)
        spec_self = AttnSpec(h, hd, causal=True, window=self.attn_window, key_bias=tgt_bias)
        spec_cross = AttnSpec(h, hd, key_bias=mem_bias)
        side = self._side_stream(mem2.device)
        kvs: List = [None] * len(self.transformer_decoder.layers)
        if side is not None:
            # kv_l = memory @ Wkv_l^T depends on the memory only: all layers' projections run beside the chain.  The
            # outputs are allocated on the current stream (they are consumed and freed there).
            cur = torch.cuda.current_stream(mem2.device)
            outs = [torch.empty((b * s, 2 * d), dtype=mem2.dtype, device=mem2.device) for _ in kvs]
            # working copies of the weights are (re)built on the current stream, which also reads them
            wkv = [c.get(layer.multihead_attn.in_proj_weight, "mat", dtype) for layer in self.transformer_decoder.layers]
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                for i, layer in enumerate(self.transformer_decoder.layers):
                    ca = layer.multihead_attn
                    ops.linear_fwd(mem2, wkv[i][d:], ca.in_proj_bias[d:], out=outs[i])
                    kvs[i] = (outs[i], side.record_event())
        for i, layer in enumerate(self.transformer_decoder.layers):
            x = self._run_layer(layer, x, mem2, b, t, s, spec_self, spec_cross, dtype, tape, training, kvs[i])
        if side is not None:
            torch.cuda.current_stream(mem2.device).wait_stream(side)
        return x

    def _run_layer(self, L: _DecoderLayer, x, mem2, b, t, s, spec_self, spec_cross, dtype, tape, training, kv_ready=None):
        c = self._wcache
        d = self.d_model
        sa, ca = L.self_attn, L.multihead_attn
        w_in = c.get(sa.in_proj_weight, "mat", dtype)
        w_o = c.get(sa.out_proj.weight, "mat", dtype)
        wc_in = c.get(ca.in_proj_weight, "mat", dtype)
        wc_o = c.get(ca.out_proj.weight, "mat", dtype)
        w1 = c.get(L.linear1.weight, "mat", dtype)
        w2 = c.get(L.linear2.weight, "mat", dtype)
        save = tape is not None
        if training and self.dropout_p > 0:
            # nn.MultiheadAttention(dropout=dropout_p) of both attention blocks: dropout on the softmax probabilities
            spec_self = ops.attn_spec_with_dropout(spec_self, self.dropout_p, self._next_seed())
            spec_cross = ops.attn_spec_with_dropout(spec_cross, self.dropout_p, self._next_seed())
        # opt-in (round-2 work): the dropouts next to the residual LayerNorms and the FFN's
        # ReLU/dropout backward folded into their neighbours (csrc/ln_fused.cu) -- 7 fewer links per layer in the chain
        fuse = training and self.dropout_p > 0 and os.environ.get("OMR_FUSE_DECODER_LINKS", "1") != "0"

        def add_ln(sub, resid, norm, seed):
            """resid + dropout(sub) -> LayerNorm; returns (y, s, stats)"""
            if seed is not None and fuse:
                return ops.dropout_add_layernorm_fwd(sub, resid, norm.weight, norm.bias, norm.eps, save, self.dropout_p, seed)
            if seed is not None:
                ops.dropout(sub, self.dropout_p, seed, inplace=True)
            return ops.add_layernorm_fwd(sub, resid, norm.weight, norm.bias, norm.eps, save)

        x2d = x.view(b * t, d)
        # --- self-attention block: x1 = LN(x + out_proj(attn(in_proj(x)))) -------------------------
        qkv = ops.linear_fwd(x2d, w_in, sa.in_proj_bias).view(b, t, 3 * d)
        o, lse = ops.attn_fwd(qkv, 0, qkv, d, qkv, 2 * d, spec_self)
        a = ops.linear_fwd(o.view(b * t, d), w_o, sa.out_proj.bias).view(b, t, d)
        seed1 = self._next_seed() if training and self.dropout_p > 0 else None
        x1, s1, st1 = add_ln(a, x, L.norm1, seed1)
        # --- cross-attention block: x2 = LN(x1 + out_proj(attn(q(x1), kv(memory)))) ----------------
        x1_2d = x1.view(b * t, d)
        q = ops.linear_fwd(x1_2d, wc_in[:d], ca.in_proj_bias[:d]).view(b, t, d)
        if kv_ready is None:
            kv = ops.linear_fwd(mem2, wc_in[d:], ca.in_proj_bias[d:]).view(b, s, 2 * d)
        else:  # projected on the side stream (see _run_stack)
            kv = kv_ready[0].view(b, s, 2 * d)
            torch.cuda.current_stream(mem2.device).wait_event(kv_ready[1])
        o2, lse2 = ops.attn_fwd(q, 0, kv, 0, kv, d, spec_cross)
        cc = ops.linear_fwd(o2.view(b * t, d), wc_o, ca.out_proj.bias).view(b, t, d)
        seed2 = self._next_seed() if training and self.dropout_p > 0 else None
        x2, s2, st2 = add_ln(cc, x1, L.norm2, seed2)
        # --- feed-forward block: x3 = LN(x2 + W2 relu(W1 x2)) -------------------------------------------
        x2_2d = x2.view(b * t, d)
        hmid = ops.linear_fwd(x2_2d, w1, L.linear1.bias, relu=True)
        ops._probe_relu(hmid)
        seed3 = self._next_seed() if training and self.dropout_p > 0 else None
        hdrop = ops.dropout(hmid, self.dropout_p, seed3) if seed3 is not None else hmid
        f = ops.linear_fwd(hdrop, w2, L.linear2.bias).view(b, t, d)
        seed4 = self._next_seed() if training and self.dropout_p > 0 else None
        x3, s3, st3 = add_ln(f, x2, L.norm3, seed4)
        if not save:
            return x3

        p = self.dropout_p
        ff = self.ff_dim
        train_w = sa.in_proj_weight.requires_grad

        def bwd(st) -> None:
            g = st["g"]  # dL/dx3 [B,T,D]
            side = st.get("side")
            cur = torch.cuda.current_stream(mem2.device) if side is not None else None

            def off_chain(fn, *tensors) -> None:
                """run fn() -- work nothing later in this chain reads -- on the side stream, after everything queued
                so far; ``tensors`` are its inputs (kept alive until the side stream is joined)"""
                if side is None:
                    fn()
                    return
                st["keep"].extend(tensors)
                side.wait_event(cur.record_event())
                with torch.cuda.stream(side):
                    fn()

            def join_if(aliased: bool) -> None:
                # without dropout d(sublayer output) IS the residual-gradient buffer, which the chain accumulates into
                if aliased and side is not None:
                    cur.wait_stream(side)

            # feed-forward block
            def ln_bwd(gy, s_, st_, norm, seed, lin_bias=None):
                """-> (ds, d sublayer output, bias gradient still to do?) of one residual block.  lin_bias: the bias of the
                linear layer that produced the sublayer output -- the fused kernel adds its gradient (column sums of the
                second output) on the way, which saves that layer's separate column-sum launch"""
                if seed is not None and fuse:
                    db = grad_buf(lin_bias) if (lin_bias is not None and train_w) else None
                    ds_, da_ = ops.layernorm_bwd_dropout(gy, s_, st_, norm.weight, grad_buf(norm.weight), grad_buf(norm.bias), p, seed,
                                                         dbias=db)
                    return ds_, da_, db is None
                ds_ = ops.layernorm_bwd(gy, s_, st_, norm.weight, grad_buf(norm.weight), grad_buf(norm.bias))
                return ds_, (ops.dropout(ds_.view(b * t, d), p, seed) if seed is not None else ds_), True

            ds3, df, need_b = ln_bwd(g, s3, st3, L.norm3, seed4, L.linear2.bias)
            ds3_2d, df = ds3.view(b * t, d), df.view(b * t, d)
            if train_w:
                gb2 = grad_buf(L.linear2.bias) if need_b else None
                off_chain(lambda: ops.linear_wgrad(hdrop, df, grad_buf(L.linear2.weight), gb2), hdrop, df)
            dh = ops.linear_dgrad(df, w2)
            if seed3 is not None and fuse:
                ops.mask_scale(dh, hdrop, 1.0 / (1.0 - p))  # zeros of hdrop = inactive or dropped
            else:
                if seed3 is not None:
                    ops.dropout(dh, p, seed3, inplace=True)
                ops.relu_bwd(hmid, dh, inplace=True)
            if train_w:
                off_chain(lambda: ops.linear_wgrad(x2_2d, dh, grad_buf(L.linear1.weight), grad_buf(L.linear1.bias)), x2_2d, dh)
            join_if(seed4 is None)
            ops.gemm(dh, w1, ds3_2d, b * t, d, ff, lda=ff, ldb=d, ldc=d, accumulate=True)  # dx2 = ds3 + dh W1
            # cross-attention block
            ds2, dcc, need_b = ln_bwd(ds3, s2, st2, L.norm2, seed2, ca.out_proj.bias)
            ds2_2d, dcc = ds2.view(b * t, d), dcc.view(b * t, d)
            if train_w:
                gbc = grad_buf(ca.out_proj.bias) if need_b else None
                off_chain(lambda: ops.linear_wgrad(o2.view(b * t, d), dcc, grad_buf(ca.out_proj.weight), gbc), o2, dcc)
            do2 = ops.linear_dgrad(dcc, wc_o).view(b, t, d)
            dq = torch.empty_like(q)
            dkv = torch.empty_like(kv)
            ops.attn_bwd(q, 0, kv, 0, kv, d, o2, do2, lse2, dq, 0, dkv, 0, dkv, d, spec_cross)
            dkv2 = dkv.view(b * s, 2 * d)
            dq2 = dq.view(b * t, d)
            if train_w:
                gw, gb = grad_buf(ca.in_proj_weight), grad_buf(ca.in_proj_bias)

                def cross_wgrads() -> None:
                    ops.linear_wgrad(mem2, dkv2, gw[d:], gb[d:])
                    ops.linear_wgrad(x1_2d, dq2, gw[:d], gb[:d])

                off_chain(cross_wgrads, mem2, dkv2, x1_2d, dq2)
            if st["need_dmem"]:
                first = st["dmem"] is None
                if first:
                    st["dmem"] = torch.empty((b, s, d), dtype=dtype, device=mem2.device)
                dmem2 = st["dmem"].view(b * s, d)
                # the memory gradient is only read after the whole stack: its 8 accumulating GEMMs stay off the chain
                off_chain(lambda: ops.gemm(dkv2, wc_in[d:], dmem2, b * s, d, 2 * d, lda=2 * d, ldb=d, ldc=d,
                                           accumulate=not first), dkv2)
            join_if(seed2 is None)
            ops.gemm(dq2, wc_in[:d], ds2_2d, b * t, d, d, lda=d, ldb=d, ldc=d, accumulate=True)  # dx1 = ds2 + dq Wq
            # self-attention block
            ds1, da, need_b = ln_bwd(ds2, s1, st1, L.norm1, seed1, sa.out_proj.bias)
            ds1_2d, da = ds1.view(b * t, d), da.view(b * t, d)
            if train_w:
                gbs = grad_buf(sa.out_proj.bias) if need_b else None
                off_chain(lambda: ops.linear_wgrad(o.view(b * t, d), da, grad_buf(sa.out_proj.weight), gbs), o, da)
            do = ops.linear_dgrad(da, w_o).view(b, t, d)
            dqkv = torch.empty_like(qkv)
            ops.attn_bwd(qkv, 0, qkv, d, qkv, 2 * d, o, do, lse, dqkv, 0, dqkv, d, dqkv, 2 * d, spec_self)
            dqkv2 = dqkv.view(b * t, 3 * d)
            if train_w:
                off_chain(lambda: ops.linear_wgrad(x2d, dqkv2, grad_buf(sa.in_proj_weight), grad_buf(sa.in_proj_bias)),
                          x2d, dqkv2)
            join_if(seed1 is None)
            ops.gemm(dqkv2, w_in, ds1_2d, b * t, d, 3 * d, lda=3 * d, ldb=d, ldc=d, accumulate=True)  # dx = ds1 + dqkv Win
            st["g"] = ds1

        tape.append(bwd)
        return x3

    def _logits_bvt(self, hidden: torch.Tensor, dtype) -> torch.Tensor:
        b, t, d = hidden.shape
        v = self.output_size
        wm = self._wcache.get(self.out_layer.weight, "mat", dtype)
        out = torch.empty((b, v, t), dtype=dtype, device=hidden.device)
        ops.gemm(wm, hidden, out, v, t, d, trans_b=True, lda=d, ldb=d, ldc=t, batch=b, stride_b=t * d, stride_c=v * t,
                 bias=self.out_layer.bias, bias_mode=2)
        return out

    def _prep(self, tgt, memory, memory_len):
        ops._lib.require_cuda(memory, "Decoder.forward")
        dtype = resolve_dtype(self.compute_dtype)
        tgt = tgt.to(memory.device)
        if tgt.shape[1] > self.max_seq_len:
            raise RuntimeError(f"target length {tgt.shape[1]} exceeds max_seq_len {self.max_seq_len}")
        mem = memory if memory.dtype == dtype and memory.is_contiguous() else None
        return dtype, tgt, mem

    def forward_hidden(self, tgt: torch.Tensor, memory: torch.Tensor, memory_len) -> torch.Tensor:
        """Everything up to (not including) the classifier: hidden [B,T,D] in the compute dtype."""
        dtype, tgt, mem = self._prep(tgt, memory, memory_len)
        if mem is None:
            mem = _CastFn.apply(memory.contiguous(), dtype)
        mem_bias, tgt_bias = self._key_biases(tgt, mem, memory_len)
        params = list(self.parameters())
        needs = torch.is_grad_enabled() and (mem.requires_grad or any(p.requires_grad for p in params))
        if needs:
            return _DecoderStackFn.apply(mem, self, tgt, mem_bias, tgt_bias, dtype, self.training, *params)
        return self._run_stack(tgt, mem, mem_bias, tgt_bias, dtype, None, self.training)

    def forward(self, tgt: torch.Tensor, memory: torch.Tensor, memory_len: Optional[torch.Tensor]) -> torch.Tensor:
        """Reference ``Decoder.forward`` (decoder.py:104-148): logits ``[B, output_size, T]``."""
        hidden = self.forward_hidden(tgt, memory, memory_len)
        dtype = hidden.dtype
        if hidden.requires_grad:
            return _LogitsFn.apply(hidden, self, dtype, self.out_layer.weight, self.out_layer.bias)
        return self._logits_bvt(hidden, dtype)

    def loss(self, tgt, memory, memory_len, targets, ignore_index: Optional[int] = None) -> torch.Tensor:
        """Fused training path: classifier + softmax cross-entropy without the [B,V,T] transpose."""
        hidden = self.forward_hidden(tgt, memory, memory_len)
        ig = self.padding_idx if ignore_index is None else ignore_index
        targets = targets.to(hidden.device)
        if hidden.requires_grad:
            return _ProjCEFn.apply(hidden, self, targets, ig, hidden.dtype, self.out_layer.weight, self.out_layer.bias)
        b, t, d = hidden.shape
        wm = self._wcache.get(self.out_layer.weight, "mat", hidden.dtype)
        if os.environ.get("OMR_FUSE_PROJ_CE", "1") != "0" and ops.proj_ce_supported(hidden.dtype, d):
            loss_out, _ = ops.proj_ce_fwd(hidden.reshape(b * t, d), wm, self.out_layer.bias, targets.reshape(-1).contiguous(), ig)
            return loss_out[0].clone()
        logits = ops.linear_fwd(hidden.reshape(b * t, d), wm, self.out_layer.bias,
                                out=ops.padded_rows(b * t, wm.shape[0], hidden.dtype, hidden.device))
        loss_out, _ = ops.ce_fwd(logits, targets.reshape(-1).contiguous(), ig)
        return loss_out[0].clone()


class _CastFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, dtype):
        ctx.src = x.dtype
        return ops.cast(x, dtype)

    @staticmethod
    def backward(ctx, g):
        return ops.cast(g.contiguous(), ctx.src), None
