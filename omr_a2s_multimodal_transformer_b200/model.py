"""Model surface: ``PositionalEncoding2D``, ``Transformer``, ``CrossAttention``, ``MultimodalTransformer``.

Drop-in for the reference ``src/transformer/model.py`` -- same constructor arguments, attribute and
sub-module names, state-dict keys, Lightning hooks and return conventions -- with every operator on
the path executed by ``libomr_b200.so``:

* encoders -> 2-D PE -> flatten/concat are fused into one "memory" producer that writes the
  ``[B, L_i + L_a, 256]`` decoder memory directly (no NCHW->NLC transpose copy, no ``torch.cat``);
* the concat mixer's key-padding mask is built by a kernel from ``xli`` / ``xla`` (no per-sample host loop);
* ``training_step`` uses the classifier + cross-entropy path without materialising ``[B,V,T]``;
* ``validation_step`` / ``test_step`` / ``get_pred_seq_and_pred_prob_seq`` run the KV-cached greedy
  decoder (``greedy.py``); ``greedy_decode_batch`` exposes the batched form.
"""
from __future__ import annotations

import math
import random
from typing import Dict, List, Optional, Tuple

import os

import torch
import torch.nn as nn

from . import ops
from .decoder import Decoder, _CastFn
from .encoder import HEIGHT_REDUCTION, WIDTH_REDUCTION, Encoder
from .greedy import BatchedGreedyDecoder
from .lightning_compat import LightningModule
from .ops import AttnSpec
from .optim import FusedAdam
from .params import MHAParams, WeightCache, grad_buf, resolve_dtype

SOS_TOKEN = "<sos>"  # reference src/data/ar_dataset.py:22
EOS_TOKEN = "<eos>"  # reference src/data/ar_dataset.py:23
NUM_CHANNELS = 1  # reference src/data/preprocessing.py:12


# ------------------------------------------------------------------------------------------------
# 2-D positional encoding and the fused memory producer
# ------------------------------------------------------------------------------------------------
class PositionalEncoding2D(nn.Module):
    """Buffer ``pe [1,C,Hmax,Wmax]`` exactly as the reference builds it (model.py:18-48)."""

    def __init__(self, num_channels: int, max_height: int, max_width: int, dropout_p: float = 0.1) -> None:
        super().__init__()
        self.dropout_p = dropout_p
        pos_h = torch.arange(max_height).unsqueeze(1)
        pos_w = torch.arange(max_width).unsqueeze(1)
        den = torch.pow(10000, torch.arange(0, num_channels // 2, 2) / num_channels)
        pe = torch.zeros(1, max_height, max_width, num_channels)
        pe[0, :, :, 0 : num_channels // 2 : 2] = torch.sin(pos_w / den).unsqueeze(0).repeat(max_height, 1, 1)
        pe[0, :, :, 1 : num_channels // 2 : 2] = torch.cos(pos_w / den).unsqueeze(0).repeat(max_height, 1, 1)
        pe[0, :, :, num_channels // 2 :: 2] = torch.sin(pos_h / den).unsqueeze(1).repeat(1, max_width, 1)
        pe[0, :, :, (num_channels // 2) + 1 :: 2] = torch.cos(pos_h / den).unsqueeze(1).repeat(1, max_width, 1)
        self.register_buffer("pe", pe.permute(0, 3, 1, 2).contiguous())
        self._nhwc: Optional[torch.Tensor] = None
        self._nhwc_key = None
        self._seed_state = 0x2545F49

    def pe_nhwc(self) -> torch.Tensor:
        """fp32 ``[Hmax,Wmax,C]`` copy of the buffer in the kernels' layout (refreshed if ``pe`` changes)."""
        key = (self.pe.data_ptr(), self.pe._version, self.pe.device)
        if self._nhwc is None or self._nhwc_key != key:
            self._nhwc = self.pe[0].permute(1, 2, 0).contiguous().float()
            self._nhwc_key = key
        return self._nhwc

    def _next_seed(self) -> int:
        self._seed_state = (self._seed_state * 6364136223846793005 + 1442695040888963407) & 0xFFFFFFFFFFFFFFFF
        return (self._seed_state >> 17) & 0x7FFFFFFF

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x ``[B,C,h,w]`` -> ``x + pe[:, :, :h, :w]`` (dropout in training), same shape."""
        b, ch, h, w = x.shape
        x_nhwc = x.permute(0, 2, 3, 1).contiguous()  # zero-copy for the encoders' channels-last outputs
        mem = _MemoryFn.apply(x_nhwc, None, self, None, self.training)
        return mem.view(b, h, w, ch).permute(0, 3, 1, 2)


class _MemoryFn(torch.autograd.Function):
    """(features_a [B,h,w,C], features_b or None) -> memory [B, L_a (+ L_b), C] = features + 2-D PE,
    segments written back to back (PE add + flatten/permute + ``torch.cat`` of model.py:495-506,654)."""

    @staticmethod
    def forward(ctx, fa: torch.Tensor, fb: Optional[torch.Tensor], pos_a: PositionalEncoding2D,
                pos_b: Optional[PositionalEncoding2D], training: bool):
        b, ha, wa, c = fa.shape
        la = ha * wa
        lb = 0 if fb is None else fb.shape[1] * fb.shape[2]
        out = torch.empty((b, la + lb, c), dtype=fa.dtype, device=fa.device)
        ops.pe2d_add(fa.contiguous(), pos_a.pe_nhwc(), out, 0)
        if fb is not None:
            ops.pe2d_add(fb.contiguous(), pos_b.pe_nhwc(), out, la)
        ctx.seeds = None
        if training and pos_a.dropout_p > 0:
            # nn.Dropout(p) of each PositionalEncoding2D; one seeded mask over the fused buffer is the
            # same distribution (independent Bernoulli per element)
            seed = pos_a._next_seed()
            ops.dropout(out, pos_a.dropout_p, seed, inplace=True)
            ctx.seeds = (pos_a.dropout_p, seed)
        ctx.shapes = (fa.shape, None if fb is None else fb.shape, la, lb)
        return out

    @staticmethod
    def backward(ctx, g: torch.Tensor):
        sa, sb, la, lb = ctx.shapes
        g = g.contiguous()
        if ctx.seeds is not None:
            g = ops.dropout(g, ctx.seeds[0], ctx.seeds[1])
        if sb is None:
            return g.view(sa), None, None, None, None
        ga = ops.copy_rows(g, 0, la).view(sa)
        gb = ops.copy_rows(g, la, lb).view(sb)
        return ga, gb, None, None, None


def _concat_key_bias(xli: torch.Tensor, xla: torch.Tensor, la_img: int, la_aud: int, device) -> torch.Tensor:
    """Key bias of mixer_concat's bool mask (model.py:660-672): -inf at image positions >= xli[b] and at
    audio positions >= xla[b] (flat-prefix semantics, as the reference)."""
    b = xli.shape[0]
    bias = torch.empty((b, la_img + la_aud), dtype=torch.float32, device=device)
    ops.key_bias_from_lengths(bias, xli.to(device=device, dtype=torch.int32).contiguous(), 0, la_img, float("-inf"))
    ops.key_bias_from_lengths(bias, xla.to(device=device, dtype=torch.int32).contiguous(), la_img, la_aud, float("-inf"))
    return bias


def _bool_prefix_mask(lens: torch.Tensor, length: int, device) -> torch.Tensor:
    pos = torch.arange(length, device=device).unsqueeze(0)
    return pos >= lens.to(device).unsqueeze(1)


# ------------------------------------------------------------------------------------------------
# Shared model behaviour
# ------------------------------------------------------------------------------------------------
class _ModelBase(LightningModule):
    def _init_common(self, w2i, i2w, ytest_i2w, max_seq_len, attn_window, teacher_forcing_prob):
        self.w2i = w2i
        self.i2w = i2w
        self.ytest_i2w = ytest_i2w if ytest_i2w is not None else i2w
        self.padding_idx = w2i["<PAD>"]
        self.max_seq_len = max_seq_len
        self.teacher_forcing_prob = teacher_forcing_prob
        self.attn_window = attn_window
        self.Y: List[List[str]] = []
        self.YHat: List[List[str]] = []
        self._compute_dtype: Optional[torch.dtype] = None
        self._greedy: Optional[BatchedGreedyDecoder] = None

    # kernel storage type: None = automatic (bf16 under autocast / OMR_COMPUTE_DTYPE, else fp32)
    def set_compute_dtype(self, dtype: Optional[torch.dtype]) -> "_ModelBase":
        self._compute_dtype = dtype
        for m in self.modules():
            if isinstance(m, (Encoder, Decoder, CrossAttention)):
                m.compute_dtype = dtype
        self._greedy = None
        return self

    @property
    def compute_dtype(self) -> torch.dtype:
        return resolve_dtype(self._compute_dtype)

    def _i2w(self, tok: int) -> str:
        w = self.i2w.get(tok)
        return w if w is not None else self.i2w[str(tok)]  # json round-trips turn int keys into str

    def _i2w_test(self, tok: int) -> str:
        w = self.ytest_i2w.get(tok)
        return w if w is not None else self.ytest_i2w[str(tok)]

    def _decoder_runner(self) -> BatchedGreedyDecoder:
        if self._greedy is None or self._greedy.dtype != self.compute_dtype:
            self._greedy = BatchedGreedyDecoder(self.decoder, self.compute_dtype)
        return self._greedy

    @torch.no_grad()
    def greedy_decode_memory(self, memory: torch.Tensor, max_steps: Optional[int] = None, stop_at_eos: bool = True,
                             use_graph: bool = True, memory_len=None):
        """memory [B,S,D] -> (tokens [B,steps], top logits [B,steps], lengths [B]) on the device.  ``memory_len`` (lengths,
        bool mask or key bias, as ``Decoder.forward`` takes it; None in the reference's batch-1 inference) masks the padded
        part of a ragged batch exactly as the teacher-forced forward does."""
        bias = None
        if memory_len is not None:
            dummy = torch.ones((memory.shape[0], 1), dtype=torch.long, device=memory.device)
            bias, _ = self.decoder._key_biases(dummy, memory, memory_len)
        return self._decoder_runner().decode(memory, self.w2i[SOS_TOKEN], self.w2i[EOS_TOKEN], self.padding_idx,
                                             max_steps=max_steps, stop_at_eos=stop_at_eos, use_graph=use_graph, mem_bias=bias)

    def _record_prediction(self, memory: torch.Tensor, y: torch.Tensor) -> None:
        toks, vals, lens = self.greedy_decode_memory(memory)
        seqs, _ = BatchedGreedyDecoder.to_lists(toks, vals, lens)
        self.YHat.append([self._i2w(t) for t in seqs[0]])
        self.Y.append([self._i2w_test(int(i)) for i in y[0][1:].tolist()])  # remove SOS_TOKEN

    @torch.no_grad()
    def on_validation_epoch_end(self, name: str = "val", print_random_samples: bool = False) -> Dict[str, float]:
        """reference model.py:205-220 / 623-636; the Levenshtein distances of all (truth, prediction) pairs of the epoch
        are computed by one launch of ``omr_levenshtein`` (``staging.compute_ed_metrics``; same rational numbers as the
        reference's ``compute_ed_metrics``, src/utils/metrics.py:52-88)"""
        from .staging import compute_ed_metrics

        metrics = compute_ed_metrics(self.Y, self.YHat, device=self.device)
        for k, v in metrics.items():
            self.log(f"{name}_{k}", v, prog_bar=True, logger=True, on_epoch=True)
        if print_random_samples:
            index = random.randint(0, len(self.Y) - 1)
            print(f"Ground truth - {self.Y[index]}")
            print(f"Prediction - {self.YHat[index]}")
        self.Y.clear()
        self.YHat.clear()
        return metrics

    @torch.no_grad()
    def on_test_epoch_end(self) -> Dict[str, float]:
        return self.on_validation_epoch_end(name="test", print_random_samples=True)

    @torch.no_grad()
    def test_step(self, batch, batch_idx) -> None:
        self.validation_step(batch, batch_idx)

    def _optimizer_params(self) -> List[nn.Parameter]:
        raise NotImplementedError

    def configure_optimizers(self):
        """Adam(lr=1e-4, amsgrad=False) over encoders + decoder (+ cross_attn), reference model.py:134-139,475-483;
        executed as one fused multi-tensor kernel."""
        opt = FusedAdam(self._optimizer_params(), lr=1e-4)
        caches = [m._wcache for m in self.modules() if isinstance(m, (Encoder, Decoder, CrossAttention))]

        def shadows(p):
            out = []
            for c in caches:
                out += c.shadows(p)
            return out

        def mark_fresh(p):
            for c in caches:
                c.mark_fresh(p)

        opt.register_shadow_provider(shadows, mark_fresh)
        return opt

    def summary(self) -> None:
        for name, mod in self.named_children():
            n = sum(p.numel() for p in mod.parameters())
            if n:
                print(f"{name}: {n:,} parameters")


# ------------------------------------------------------------------------------------------------
# Unimodal transformer (reference model.py:54-262)
# ------------------------------------------------------------------------------------------------
class Transformer(_ModelBase):
    def __init__(self, max_input_height: int, max_input_width: int, max_seq_len: int, w2i: Dict[str, int],
                 i2w: Dict[int, str], ytest_i2w: Optional[Dict[int, str]] = None, attn_window: int = -1,
                 teacher_forcing_prob: float = 0.5) -> None:
        super().__init__()
        self.save_hyperparameters()
        self._init_common(w2i, i2w, ytest_i2w, max_seq_len, attn_window, teacher_forcing_prob)
        self.max_input_height = max_input_height
        self.max_input_width = max_input_width
        self.encoder = Encoder(in_channels=NUM_CHANNELS)
        self.pos_2d = PositionalEncoding2D(
            num_channels=256,
            max_height=math.ceil(max_input_height / HEIGHT_REDUCTION),
            max_width=math.ceil(max_input_width / WIDTH_REDUCTION),
        )
        self.decoder = Decoder(output_size=len(w2i), max_seq_len=max_seq_len, num_embeddings=len(w2i),
                               padding_idx=self.padding_idx, attn_window=attn_window)
        self.compute_loss = nn.CrossEntropyLoss(ignore_index=self.padding_idx)

    def _optimizer_params(self):
        return list(self.encoder.parameters()) + list(self.decoder.parameters())

    def encode(self, x: torch.Tensor) -> torch.Tensor:
        """x [B,1,H,W] -> decoder memory [B, h*w, 256] (encoder + 2-D PE + flatten/permute, model.py:141-147)."""
        f = self.encoder.forward_nhwc(x)
        return _MemoryFn.apply(f, None, self.pos_2d, None, self.training)

    def forward(self, x: torch.Tensor, xl: Optional[torch.Tensor], y_in: torch.Tensor) -> torch.Tensor:
        return self.decoder(tgt=y_in, memory=self.encode(x), memory_len=xl)

    def apply_teacher_forcing(self, y: torch.Tensor) -> torch.Tensor:
        """reference model.py:152-160: each non-pad token is replaced with probability ``teacher_forcing_prob`` by
        ``randint(0, V-1)``.  The reference draws token by token in a B x T Python loop that synchronises with the device
        at every element; the same distribution is drawn here on the device in one pass (as the reference's own
        multimodal variant does, model.py:545-559), which also keeps the step capturable in a CUDA graph."""
        random_mask = torch.rand_like(y, dtype=torch.float) < self.teacher_forcing_prob
        combined = random_mask & (y != self.padding_idx)
        random_indices = torch.randint(0, len(self.w2i), y.shape, device=y.device)
        return torch.where(combined, random_indices, y)

    def training_step(self, batch, batch_idx) -> torch.Tensor:
        x, xl, y_in, y_out = batch
        y_in = self.apply_teacher_forcing(y_in)
        loss = self.decoder.loss(tgt=y_in, memory=self.encode(x), memory_len=xl, targets=y_out)
        self.log("train_loss", loss, prog_bar=True, logger=True, on_epoch=True)
        return loss

    @torch.no_grad()
    def validation_step(self, batch, batch_idx) -> None:
        x, y = batch
        assert x.size(0) == y.size(0) == 1, "Inference only supports batch_size = 1"
        self._record_prediction(self.encode(x), y)

    @torch.no_grad()
    def get_pred_seq_and_pred_prob_seq(self, x: torch.Tensor) -> Tuple[List[str], List[float]]:
        """reference model.py:226-262: predicted words and the raw top logit of every step."""
        assert x.size(0) == 1, "Inference only supports batch_size = 1"
        toks, vals, lens = self.greedy_decode_memory(self.encode(x))
        seqs, probs = BatchedGreedyDecoder.to_lists(toks, vals, lens)
        return [self._i2w(t) for t in seqs[0]], probs[0]

    @torch.no_grad()
    def greedy_decode_batch(self, x: torch.Tensor, max_steps: Optional[int] = None, stop_at_eos: bool = True):
        return self.greedy_decode_memory(self.encode(x), max_steps=max_steps, stop_at_eos=stop_at_eos)


# ------------------------------------------------------------------------------------------------
# Cross-attention mixer (reference model.py:268-355)
# ------------------------------------------------------------------------------------------------
class _CrossAttnFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, query, key_value, mod: "CrossAttention", q_len, kv_len, dtype, *params):
        att = mod.attention
        c = mod._wcache
        b, tq, d = query.shape
        tk = key_value.shape[1]
        w_in = c.get(att.in_proj_weight, "mat", dtype)
        w_o = c.get(att.out_proj.weight, "mat", dtype)
        q2, kv2 = query.reshape(b * tq, d), key_value.reshape(b * tk, d)
        q = ops.linear_fwd(q2, w_in[:d], att.in_proj_bias[:d]).view(b, tq, d)
        kv = ops.linear_fwd(kv2, w_in[d:], att.in_proj_bias[d:]).view(b, tk, 2 * d)
        spec = AttnSpec(att.num_heads, att.head_dim, q_len=q_len, kv_len=kv_len, quirk_mod=b if q_len is not None else 0)
        if mod.training and mod.dropout_p > 0:  # nn.MultiheadAttention(dropout=0.1) of the mixer (reference model.py:292-297)
            mod._drop_seed = (mod._drop_seed * 6364136223846793005 + 1442695040888963407) & 0xFFFFFFFFFFFFFFFF
            spec = ops.attn_spec_with_dropout(spec, mod.dropout_p, (mod._drop_seed >> 17) & 0x7FFFFFFF)
        o, lse = ops.attn_fwd(q, 0, kv, 0, kv, d, spec)
        out = ops.linear_fwd(o.view(b * tq, d), w_o, att.out_proj.bias).view(b, tq, d)
        ctx.mod, ctx.dtype, ctx.spec = mod, dtype, spec
        ctx.save_for_backward(q2, kv2, q, kv, o, lse)
        return out

    @staticmethod
    def backward(ctx, g):
        q2, kv2, q, kv, o, lse = ctx.saved_tensors
        mod, dtype, spec = ctx.mod, ctx.dtype, ctx.spec
        att, c = mod.attention, mod._wcache
        b, tq, d = q.shape
        tk = kv.shape[1]
        w_in = c.get(att.in_proj_weight, "mat", dtype)
        w_o = c.get(att.out_proj.weight, "mat", dtype)
        g2 = g.contiguous().view(b * tq, d)
        train_w = att.in_proj_weight.requires_grad
        if train_w:
            ops.linear_wgrad(o.view(b * tq, d), g2, grad_buf(att.out_proj.weight), grad_buf(att.out_proj.bias))
        do = ops.linear_dgrad(g2, w_o).view(b, tq, d)
        dq, dkv = torch.empty_like(q), torch.empty_like(kv)
        ops.attn_bwd(q, 0, kv, 0, kv, d, o, do, lse, dq, 0, dkv, 0, dkv, d, spec)
        dq2, dkv2 = dq.view(b * tq, d), dkv.view(b * tk, 2 * d)
        if train_w:
            gw, gb = grad_buf(att.in_proj_weight), grad_buf(att.in_proj_bias)
            ops.linear_wgrad(q2, dq2, gw[:d], gb[:d])
            ops.linear_wgrad(kv2, dkv2, gw[d:], gb[d:])
        dquery = ops.linear_dgrad(dq2, w_in[:d]).view(b, tq, d) if ctx.needs_input_grad[0] else None
        dkeyval = ops.linear_dgrad(dkv2, w_in[d:]).view(b, tk, d) if ctx.needs_input_grad[1] else None
        return (dquery, dkeyval, None, None, None, None) + tuple(None for _ in ctx.needs_input_grad[6:])


class CrossAttention(nn.Module):
    """nn.MultiheadAttention(feature_dim, num_heads, batch_first) with query = one modality and
    key = value = the other; the reference's block mask (rows >= len_query AND cols >= len_key_value)
    is reproduced including its head-major ``repeat`` ordering.  The
    head-averaged attention weights the reference also returns are never used by its callers
    (model.py:687,704) and are not computed; ``forward`` returns ``(attn_output, None)``."""

    def __init__(self, feature_dim: int, num_heads: int = 4, dropout: float = 0.1) -> None:
        super().__init__()
        self.num_heads = num_heads
        self.dropout_p = dropout
        self.attention = MHAParams(feature_dim, num_heads)
        self.compute_dtype: Optional[torch.dtype] = None
        self._wcache = WeightCache()
        self._drop_seed = 0x1357911

    def forward(self, query, len_query, key_value, len_key_value):
        dtype = resolve_dtype(self.compute_dtype)
        ops._lib.require_cuda(query, "CrossAttention.forward")
        q_len = kv_len = None
        if len_query is not None and len_key_value is not None:
            q_len = len_query.to(device=query.device, dtype=torch.int32).contiguous()
            kv_len = len_key_value.to(device=query.device, dtype=torch.int32).contiguous()
        query, key_value = query.contiguous(), key_value.contiguous()
        if query.dtype != dtype:
            query = _CastFn.apply(query, dtype)
        if key_value.dtype != dtype:
            key_value = _CastFn.apply(key_value, dtype)
        params = list(self.parameters())
        return _CrossAttnFn.apply(query, key_value, self, q_len, kv_len, dtype, *params), None


# ------------------------------------------------------------------------------------------------
# Multimodal transformer (reference model.py:358-726)
# ------------------------------------------------------------------------------------------------
class MultimodalTransformer(_ModelBase):
    def __init__(self, max_img_height: int, max_img_width: int, max_audio_height: int, max_audio_width: int,
                 max_seq_len: int, w2i: Dict[str, int], i2w: Dict[int, str], ytest_i2w: Optional[Dict[int, str]] = None,
                 mixer_type: str = "concat", attn_window: int = -1, teacher_forcing_prob: float = 0.5,
                 teacher_forcing_modality_prob: float = 0.5) -> None:
        super().__init__()
        self.save_hyperparameters()
        self._init_common(w2i, i2w, ytest_i2w, max_seq_len, attn_window, teacher_forcing_prob)
        self.max_img_height, self.max_img_width = max_img_height, max_img_width
        self.max_audio_height, self.max_audio_width = max_audio_height, max_audio_width
        self.teacher_forcing_modality_prob = teacher_forcing_modality_prob
        self.mixer_type = mixer_type
        self.image_encoder = Encoder(in_channels=NUM_CHANNELS)
        self.image_pos_2d = PositionalEncoding2D(256, math.ceil(max_img_height / HEIGHT_REDUCTION),
                                                 math.ceil(max_img_width / WIDTH_REDUCTION))
        self.audio_encoder = Encoder(in_channels=NUM_CHANNELS)
        self.audio_pos_2d = PositionalEncoding2D(256, math.ceil(max_audio_height / HEIGHT_REDUCTION),
                                                 math.ceil(max_audio_width / WIDTH_REDUCTION))
        self.decoder = Decoder(output_size=len(w2i), max_seq_len=max_seq_len, num_embeddings=len(w2i),
                               padding_idx=self.padding_idx, attn_window=attn_window)
        if mixer_type == "concat":
            self.mixer = self.mixer_concat
        elif mixer_type == "attn_img":
            self.cross_attn = CrossAttention(feature_dim=256)
            self.mixer = self.mixer_attn_img
        elif mixer_type == "attn_audio":
            self.cross_attn = CrossAttention(feature_dim=256)
            self.mixer = self.mixer_attn_audio
        elif mixer_type == "attn_both":
            self.cross_attn = CrossAttention(feature_dim=256)
            self.mixer = self.mixer_attn_both
        else:
            raise ValueError(f"Invalid mixer type: {mixer_type}")
        self.compute_loss = nn.CrossEntropyLoss(ignore_index=self.padding_idx)

    def _optimizer_params(self):
        params = (list(self.image_encoder.parameters()) + list(self.audio_encoder.parameters())
                  + list(self.decoder.parameters()))
        if hasattr(self, "cross_attn"):
            params += list(self.cross_attn.parameters())
        return params

    # ---- encoders + fusion --------------------------------------------------------------------
    def _memory(self, xi, xa, xli, xla, modality: str):
        """-> (memory [B,S,256], memory_len for the decoder).  For the concat mixer the fused memory is
        written in one pass and the mask is a kernel-built fp32 key bias (-inf on padded frames)."""
        fi, fa = self._encode_both(xi, xa)
        if modality == "image":  # model.py:510-512
            return _MemoryFn.apply(fi, None, self.image_pos_2d, None, self.training), xli
        if modality == "audio":  # model.py:513-515
            return _MemoryFn.apply(fa, None, self.audio_pos_2d, None, self.training), xla
        if modality != "both":
            raise ValueError(f"Invalid modality: {modality}")
        if self.mixer_type == "concat":
            mem = _MemoryFn.apply(fi, fa, self.image_pos_2d, self.audio_pos_2d, self.training)
            li, la = fi.shape[1] * fi.shape[2], fa.shape[1] * fa.shape[2]
            xl = None if (xli is None or xla is None) else _concat_key_bias(xli, xla, li, la, mem.device)
            return mem, xl
        mi = _MemoryFn.apply(fi, None, self.image_pos_2d, None, self.training)
        ma = _MemoryFn.apply(fa, None, self.audio_pos_2d, None, self.training)
        return self.mixer(xi=mi, xa=ma, xli=xli, xla=xla)

    def _encode_both(self, xi, xa):
        """Both encoders always run (reference model.py:485-506) and are independent: the audio encoder is issued on a
        side stream so that its (many small, latency-bound) kernels fill the gaps of the image encoder's.  Each encoder
        is one autograd node, and autograd replays a node's backward on the stream of its forward, so the two backward
        passes overlap the same way.  ``OMR_OVERLAP_ENCODERS=0`` keeps everything on one stream."""
        if os.environ.get("OMR_OVERLAP_ENCODERS", "1") == "0" or not xi.is_cuda:
            return self.image_encoder.forward_nhwc(xi), self.audio_encoder.forward_nhwc(xa)
        cur = torch.cuda.current_stream(xi.device)
        side = getattr(self, "_enc_side_stream", None)
        if side is None or side.device != xi.device:
            # as important as the stream it forks from (see graph.GraphedTrainStep): high priority
            prio = os.environ.get("OMR_STREAM_PRIORITY", "1") != "0"
            side = torch.cuda.Stream(device=xi.device, priority=-1) if prio else torch.cuda.Stream(device=xi.device)
            self._enc_side_stream = side
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            fa = self.audio_encoder.forward_nhwc(xa)
        fi = self.image_encoder.forward_nhwc(xi)
        cur.wait_stream(side)
        fa.record_stream(cur)
        return fi, fa

    def encoder_forward(self, xi, xa, xli=None, xla=None, apply_teacher_forcing_modality: bool = False):
        """reference model.py:485-522: returns (memory, xl) with xl a bool [B,S] mask (concat), the
        chosen modality's lengths, or None."""
        modality = self.apply_teacher_forcing_modality() if apply_teacher_forcing_modality else "both"
        mem, xl = self._memory(xi, xa, xli, xla, modality)
        if modality == "both" and self.mixer_type == "concat" and xl is not None:
            xl = torch.isinf(xl)  # the reference's bool key-padding mask
        return mem, xl

    def forward(self, xi, xli, xa, xla, y_in, apply_teacher_forcing_modality: bool = False) -> torch.Tensor:
        modality = self.apply_teacher_forcing_modality() if apply_teacher_forcing_modality else "both"
        mem, xl = self._memory(xi, xa, xli, xla, modality)
        return self.decoder(tgt=y_in, memory=mem, memory_len=xl)

    def apply_teacher_forcing(self, y: torch.Tensor) -> torch.Tensor:
        """reference model.py:545-559"""
        random_mask = torch.rand_like(y, dtype=torch.float) < self.teacher_forcing_prob
        combined = random_mask & (y != self.padding_idx)
        random_indices = torch.randint(0, len(self.w2i), y.shape, device=y.device)
        return torch.where(combined, random_indices, y)

    def apply_teacher_forcing_modality(self) -> str:
        """reference model.py:561-575"""
        if random.random() < self.teacher_forcing_modality_prob:
            return "image" if random.random() < 0.5 else "audio"
        return "both"

    def training_step(self, batch, batch_idx) -> torch.Tensor:
        xi, xli, xa, xla, y_in, y_out = batch
        y_in = self.apply_teacher_forcing(y_in)
        modality = self.apply_teacher_forcing_modality()
        mem, xl = self._memory(xi, xa, xli, xla, modality)
        loss = self.decoder.loss(tgt=y_in, memory=mem, memory_len=xl, targets=y_out)
        self.log("train_loss", loss, prog_bar=True, logger=True, on_epoch=True)
        return loss

    @torch.no_grad()
    def validation_step(self, batch, batch_idx) -> None:
        xi, xa, y = batch
        assert xi.size(0) == xa.size(0) == y.size(0) == 1, "Inference only supports batch_size = 1"
        mem, _ = self._memory(xi, xa, None, None, "both")
        self._record_prediction(mem, y)

    @torch.no_grad()
    def greedy_decode_batch(self, xi, xa, max_steps: Optional[int] = None, stop_at_eos: bool = True, xli=None, xla=None):
        """batched KV-cached greedy decoding; xli / xla (frame counts of a padded batch, as ``forward`` takes them) mask the
        padded memory positions the way the teacher-forced forward does"""
        mem, xl = self._memory(xi, xa, xli, xla, "both")
        return self.greedy_decode_memory(mem, max_steps=max_steps, stop_at_eos=stop_at_eos, memory_len=xl)

    # ---- modality mixers (reference model.py:644-726) -----------------------------------------------
    def mixer_concat(self, xi, xa, xli=None, xla=None):
        x = torch.cat([xi, xa], dim=1)
        if xli is not None and xla is not None:
            xl = torch.cat([_bool_prefix_mask(xli, xi.shape[1], xi.device), _bool_prefix_mask(xla, xa.shape[1], xa.device)], dim=1)
        else:
            xl = None
        return x, xl

    def mixer_attn_img(self, xi, xa, xli=None, xla=None):
        x, _ = self.cross_attn(query=xa, len_query=xla, key_value=xi, len_key_value=xli)
        return x, (xla if (xli is not None and xla is not None) else None)

    def mixer_attn_audio(self, xi, xa, xli=None, xla=None):
        x, _ = self.cross_attn(query=xi, len_query=xli, key_value=xa, len_key_value=xla)
        return x, (xli if (xli is not None and xla is not None) else None)

    def mixer_attn_both(self, xi, xa, xli=None, xla=None):
        xa, xla = self.mixer_attn_img(xi=xi, xa=xa, xli=xli, xla=xla)
        xi, xli = self.mixer_attn_audio(xi=xi, xa=xa, xli=xli, xla=xla)
        return self.mixer_concat(xi=xi, xa=xa, xli=xli, xla=xla)
