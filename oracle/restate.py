"""TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference hot path in plain torch CPU ops.

Every function works on a flat ``state_dict`` (reference key names) and explicit tensors, in
fp32 or fp64 (``dtype=``), eval-mode semantics (dropout off; InstanceNorm uses instance
statistics in both modes, reference ``src/transformer/encoder.py:151-156``).  It is the checker
for the CUDA path; it is pinned against the real reference by ``tests/test_oracle_pin.py`` and
``tests/golden`` (see ``oracle/__init__.py``).

Citations are ``path:line`` relative to the reference root.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]

CONV_BLOCK_STRIDES = [(1, 1), (2, 2), (2, 2), (2, 2), (2, 1)]  # src/transformer/encoder.py:255-259
IN_EPS = 1e-3  # src/transformer/encoder.py:153
LN_EPS = 1e-5  # torch nn.TransformerDecoderLayer default (src/transformer/decoder.py:86-95)


# Test hook: when set, every ReLU of the path goes through RELU_HOOK(x) instead of F.relu(x) -- the parity tests use it to
# record pre-activations and to replay another implementation's ReLU decisions (tests/helpers.relu_decisions).
RELU_HOOK = None


def _relu(x: torch.Tensor) -> torch.Tensor:
    return F.relu(x) if RELU_HOOK is None else RELU_HOOK(x)


def _p(sd: SD, key: str, dtype: torch.dtype) -> torch.Tensor:
    return sd[key].to(dtype)


# --------------------------------------------------------------------------------------------
# Encoder -- src/transformer/encoder.py
# --------------------------------------------------------------------------------------------


def instance_norm(x: torch.Tensor, eps: float = IN_EPS) -> torch.Tensor:
    """nn.InstanceNorm2d(eps=1e-3, affine=False, track_running_stats=False): encoder.py:151-156."""
    mean = x.mean(dim=(2, 3), keepdim=True)
    var = x.var(dim=(2, 3), unbiased=False, keepdim=True)
    return (x - mean) / torch.sqrt(var + eps)


def conv_block(sd: SD, pre: str, x: torch.Tensor, stride: Tuple[int, int], drop=None) -> torch.Tensor:
    """ConvBlock.forward: encoder.py:159-181.  ``drop`` (train mode only): callable(slot, x) -> x applied at the three
    MixDropout slots (after the ReLUs of conv1, conv2, conv3); None = eval mode."""
    dt = x.dtype
    drop = drop or (lambda slot, t: t)
    x = drop(1, _relu(F.conv2d(x, _p(sd, pre + "conv1.weight", dt), _p(sd, pre + "conv1.bias", dt), padding=1)))
    x = drop(2, _relu(F.conv2d(x, _p(sd, pre + "conv2.weight", dt), _p(sd, pre + "conv2.bias", dt), padding=1)))
    x = instance_norm(x)
    x = drop(3, _relu(F.conv2d(x, _p(sd, pre + "conv3.weight", dt), _p(sd, pre + "conv3.bias", dt), padding=1, stride=stride)))
    return x


def depth_sep_conv(sd: SD, pre: str, x: torch.Tensor) -> torch.Tensor:
    """DepthSepConv2D.forward (3x3 depthwise pad 1, then 1x1 pointwise): encoder.py:56-84."""
    dt = x.dtype
    c = x.shape[1]
    x = F.conv2d(x, _p(sd, pre + "depth_conv.weight", dt), _p(sd, pre + "depth_conv.bias", dt), padding=1, groups=c)
    x = F.conv2d(x, _p(sd, pre + "point_conv.weight", dt), _p(sd, pre + "point_conv.bias", dt))
    return x


def dsc_block(sd: SD, pre: str, x: torch.Tensor, drop=None) -> torch.Tensor:
    """DSCBlock.forward (ReLU after conv1/conv2 only): encoder.py:218-238; ``drop`` as in conv_block."""
    drop = drop or (lambda slot, t: t)
    x = drop(1, _relu(depth_sep_conv(sd, pre + "conv1.", x)))
    x = drop(2, _relu(depth_sep_conv(sd, pre + "conv2.", x)))
    x = instance_norm(x)
    x = drop(3, depth_sep_conv(sd, pre + "conv3.", x))
    return x


def encoder_forward(sd: SD, pre: str, x: torch.Tensor, drop=None) -> torch.Tensor:
    """Encoder.forward: encoder.py:271-291.  x [B,1,H,W] -> [B,256,ceil(H/16),ceil(W/8)].
    ``drop``: callable(block_index 0..8, slot 1..3, x) -> x for a train-mode pass with given dropout masks."""
    bd = (lambda b: None) if drop is None else (lambda b: (lambda slot, t: drop(b, slot, t)))
    for i, s in enumerate(CONV_BLOCK_STRIDES):
        x = conv_block(sd, f"{pre}conv_blocks.{i}.", x, s, bd(i))
    for i in range(4):
        xt = dsc_block(sd, f"{pre}dscblocks.{i}.", x, bd(len(CONV_BLOCK_STRIDES) + i))
        x = x + xt if x.shape == xt.shape else xt  # encoder.py:287-289
    return x


# --------------------------------------------------------------------------------------------
# Positional encodings -- src/transformer/model.py:18-48, src/transformer/decoder.py:7-32
# --------------------------------------------------------------------------------------------


def pe2d_table(num_channels: int, max_h: int, max_w: int) -> torch.Tensor:
    """PositionalEncoding2D buffer ``pe`` [1,C,Hmax,Wmax]: model.py:31-42."""
    pos_h = torch.arange(max_h).unsqueeze(1)
    pos_w = torch.arange(max_w).unsqueeze(1)
    den = torch.pow(10000, torch.arange(0, num_channels // 2, 2) / num_channels)
    pe = torch.zeros(1, max_h, max_w, num_channels)
    pe[0, :, :, 0 : num_channels // 2 : 2] = torch.sin(pos_w / den).unsqueeze(0).repeat(max_h, 1, 1)
    pe[0, :, :, 1 : num_channels // 2 : 2] = torch.cos(pos_w / den).unsqueeze(0).repeat(max_h, 1, 1)
    pe[0, :, :, num_channels // 2 :: 2] = torch.sin(pos_h / den).unsqueeze(1).repeat(1, max_w, 1)
    pe[0, :, :, (num_channels // 2) + 1 :: 2] = torch.cos(pos_h / den).unsqueeze(1).repeat(1, max_w, 1)
    return pe.permute(0, 3, 1, 2).contiguous()


def pe1d_table(max_len: int, emb_dim: int) -> torch.Tensor:
    """PositionalEncoding1D buffer ``pe`` [1,max_len,D]: decoder.py:21-28."""
    pos = torch.arange(max_len).unsqueeze(1)
    den = torch.pow(10000, torch.arange(0, emb_dim, 2) / emb_dim)
    pe = torch.zeros(1, max_len, emb_dim)
    pe[0, :, 0::2] = torch.sin(pos / den)
    pe[0, :, 1::2] = torch.cos(pos / den)
    return pe


def encode_to_memory(sd: SD, enc_pre: str, pe_key: str, x: torch.Tensor) -> torch.Tensor:
    """encoder -> + pe[:, :, :h, :w] -> flatten(2).permute(0,2,1): model.py:141-147 / 495-506."""
    f = encoder_forward(sd, enc_pre, x)
    f = f + sd[pe_key].to(f.dtype)[:, :, : f.shape[2], : f.shape[3]]
    return f.flatten(2).permute(0, 2, 1).contiguous()


# --------------------------------------------------------------------------------------------
# Attention (torch nn.MultiheadAttention semantics used at decoder.py:86-95 and model.py:292-297)
# --------------------------------------------------------------------------------------------


def mha(
    sd: SD,
    pre: str,
    query: torch.Tensor,
    key: torch.Tensor,
    value: torch.Tensor,
    nhead: int,
    attn_mask: Optional[torch.Tensor] = None,  # float additive [Tq,Tk] or [B*nhead,Tq,Tk]; bool -> -inf
    key_padding_mask: Optional[torch.Tensor] = None,  # [B,Tk] bool (-inf) or float (ADDED AS-IS)
) -> torch.Tensor:
    """Packed in-proj (q rows 0:D, k D:2D, v 2D:3D), heads of D/nhead, scale 1/sqrt(hd), masks
    summed before softmax; float masks are additive."""
    dt = query.dtype
    B, Tq, D = query.shape
    Tk = key.shape[1]
    hd = D // nhead
    w = _p(sd, pre + "in_proj_weight", dt)
    b = _p(sd, pre + "in_proj_bias", dt)
    q = F.linear(query, w[:D], b[:D]).view(B, Tq, nhead, hd).transpose(1, 2)
    k = F.linear(key, w[D : 2 * D], b[D : 2 * D]).view(B, Tk, nhead, hd).transpose(1, 2)
    v = F.linear(value, w[2 * D :], b[2 * D :]).view(B, Tk, nhead, hd).transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
    if attn_mask is not None:
        m = attn_mask
        if m.dtype == torch.bool:
            m = torch.zeros(m.shape, dtype=dt).masked_fill(m, float("-inf"))
        if m.dim() == 3:
            m = m.view(B, nhead, Tq, Tk)
        s = s + m.to(dt)
    if key_padding_mask is not None:
        m = key_padding_mask
        if m.dtype == torch.bool:
            m = torch.zeros(m.shape, dtype=dt).masked_fill(m, float("-inf"))
        s = s + m.to(dt)[:, None, None, :]
    p = torch.softmax(s, dim=-1)
    o = (p @ v).transpose(1, 2).reshape(B, Tq, D)
    return F.linear(o, _p(sd, pre + "out_proj.weight", dt), _p(sd, pre + "out_proj.bias", dt))


# --------------------------------------------------------------------------------------------
# Decoder -- src/transformer/decoder.py
# --------------------------------------------------------------------------------------------


def window_mask(size: int, window: int, dtype: torch.dtype) -> torch.Tensor:
    """create_variable_window_mask: decoder.py:191-217 (row i sees keys max(0,i-w)..i if w<size)."""
    i = torch.arange(size).unsqueeze(1)
    j = torch.arange(size).unsqueeze(0)
    ok = j <= i
    if window < size:
        ok = ok & (j >= i - window)
    return torch.zeros(size, size, dtype=dtype).masked_fill(~ok, float("-inf"))


def tgt_masks(tgt: torch.Tensor, attn_window: int, dtype: torch.dtype):
    """get_tgt_masks: decoder.py:219-254 (causal or window; pad mask = (tgt == 0) as FLOAT)."""
    T = tgt.shape[1]
    if attn_window > 0:
        m = window_mask(T, attn_window, dtype)
    else:
        m = window_mask(T, T, dtype)  # == generate_square_subsequent_mask
    return m, (tgt == 0).to(dtype)


def memory_key_padding_mask(memory: torch.Tensor, memory_len: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    """get_memory_key_padding_mask: decoder.py:150-189.  bool -> cloned (true masking);
    integer lengths -> FLOAT 0/1 mask, which torch treats as additive (+1.0 on padded keys)."""
    if memory_len is None:
        return None
    if memory_len.dtype == torch.bool:
        assert memory_len.shape[0] == memory.shape[0] and memory_len.shape[1] == memory.shape[1]
        return memory_len.clone()
    m = torch.zeros(memory.shape[:2], dtype=memory.dtype)
    for i, l in enumerate(memory_len.tolist()):
        m[i, int(l) :] = 1
    return m


def decoder_layer(sd: SD, pre: str, x, memory, tmask, tkpm, mkpm, nhead: int, drop=None) -> torch.Tensor:
    """torch nn.TransformerDecoderLayer, post-norm, ReLU FFN (built at decoder.py:86-95).  ``drop`` (train mode with given
    masks): callable(slot, x) -> x at dropout1 (1), dropout2 (2), the FFN's inner dropout (3) and dropout3 (4); the
    attention-probability dropout is not replayed here (tests/test_gpu_ops.py covers it at kernel level)."""
    dt = x.dtype
    D = x.shape[-1]
    drop = drop or (lambda slot, t: t)

    def ln(t, name):
        return F.layer_norm(t, (D,), _p(sd, pre + name + ".weight", dt), _p(sd, pre + name + ".bias", dt), LN_EPS)

    x = ln(x + drop(1, mha(sd, pre + "self_attn.", x, x, x, nhead, attn_mask=tmask, key_padding_mask=tkpm)), "norm1")
    x = ln(x + drop(2, mha(sd, pre + "multihead_attn.", x, memory, memory, nhead, key_padding_mask=mkpm)), "norm2")
    h = drop(3, _relu(F.linear(x, _p(sd, pre + "linear1.weight", dt), _p(sd, pre + "linear1.bias", dt))))
    h = F.linear(h, _p(sd, pre + "linear2.weight", dt), _p(sd, pre + "linear2.bias", dt))
    return ln(x + drop(4, h), "norm3")


def decoder_hidden(
    sd: SD,
    pre: str,
    tgt: torch.Tensor,
    memory: torch.Tensor,
    memory_len: Optional[torch.Tensor],
    attn_window: int = -1,
    nhead: int = 4,
    num_layers: int = 8,
    drop=None,
) -> torch.Tensor:
    """Decoder.forward up to (not including) the classifier: decoder.py:104-143.
    ``drop``: callable(layer index or -1 for the embedding dropout, slot, x) -> x (train mode with given masks)."""
    dt = memory.dtype
    emb = _p(sd, pre + "embedding.weight", dt)[tgt]  # row padding_idx is zero (decoder.py:73-77)
    x = emb + sd[pre + "pos_1d.pe"].to(dt)[:, : tgt.shape[1], :]  # decoder.py:29-32
    if drop is not None:
        x = drop(-1, 0, x)
    mkpm = memory_key_padding_mask(memory, memory_len)
    tmask, tkpm = tgt_masks(tgt, attn_window, dt)
    tkpm = None if mkpm is None else tkpm  # decoder.py:132
    for i in range(num_layers):
        ld = None if drop is None else (lambda slot, t, i=i: drop(i, slot, t))
        x = decoder_layer(sd, f"{pre}transformer_decoder.layers.{i}.", x, memory, tmask, tkpm, mkpm, nhead, ld)
    return x


def decoder_forward(sd: SD, pre: str, tgt, memory, memory_len, attn_window: int = -1, nhead: int = 4, num_layers: int = 8,
                    drop=None):
    """Decoder.forward: decoder.py:104-148 -> logits [B,V,T] (Conv1d k=1 on the permuted hidden)."""
    dt = memory.dtype
    h = decoder_hidden(sd, pre, tgt, memory, memory_len, attn_window, nhead, num_layers, drop)
    w = _p(sd, pre + "out_layer.weight", dt)  # [V,D,1]
    return F.conv1d(h.permute(0, 2, 1).contiguous(), w, _p(sd, pre + "out_layer.bias", dt))


def ce_loss(logits_bvt: torch.Tensor, y_out: torch.Tensor, ignore_index: int = 0) -> torch.Tensor:
    """CrossEntropyLoss(ignore_index=pad) on [B,V,T] logits: model.py:109,166,444,588."""
    return F.cross_entropy(logits_bvt, y_out, ignore_index=ignore_index)


# --------------------------------------------------------------------------------------------
# Mixers -- src/transformer/model.py:268-355, 644-726
# --------------------------------------------------------------------------------------------


def mixer_concat(xi, xa, xli=None, xla=None):
    """mixer_concat: model.py:644-675 (bool key-padding mask, flat-prefix semantics per segment)."""
    x = torch.cat([xi, xa], dim=1)
    if xli is None or xla is None:
        return x, None
    mi = torch.zeros(xi.shape[:2], dtype=torch.bool)
    for i, l in enumerate(xli.tolist()):
        mi[i, int(l) :] = True
    ma = torch.zeros(xa.shape[:2], dtype=torch.bool)
    for i, l in enumerate(xla.tolist()):
        ma[i, int(l) :] = True
    return x, torch.cat([mi, ma], dim=1)


def cross_attention(sd: SD, pre: str, query, len_query, key_value, len_key_value, nhead: int = 4):
    """CrossAttention.forward: model.py:299-355.  The bool mask [B,Tq,Tk] (True on rows>=lq AND
    cols>=lkv) is ``repeat(num_heads,1,1)``-ed, i.e. tiled HEAD-major, while torch MHA indexes the
    flattened [B*nhead] axis BATCH-major: entry (b,h) therefore uses sample (b*nhead+h) % B's mask
    (SURVEY.md section 8 a16).  Reproduced, not fixed."""
    B, Tq, _ = query.shape
    Tk = key_value.shape[1]
    attn_mask = None
    if len_query is not None and len_key_value is not None:
        m = torch.zeros(B, Tq, Tk, dtype=torch.bool)
        for i, (lq, lkv) in enumerate(zip(len_query.tolist(), len_key_value.tolist())):
            m[i, int(lq) :, int(lkv) :] = True
        attn_mask = m.repeat(nhead, 1, 1)
    return mha(sd, pre + "attention.", query, key_value, key_value, nhead, attn_mask=attn_mask)


def mix(sd: SD, mixer_type: str, xi, xa, xli=None, xla=None):
    both = xli is not None and xla is not None
    if mixer_type == "concat":
        return mixer_concat(xi, xa, xli, xla)
    if mixer_type == "attn_img":  # model.py:677-692: audio queries attend image keys
        return cross_attention(sd, "cross_attn.", xa, xla, xi, xli), (xla if both else None)
    if mixer_type == "attn_audio":  # model.py:694-709
        return cross_attention(sd, "cross_attn.", xi, xli, xa, xla), (xli if both else None)
    if mixer_type == "attn_both":  # model.py:711-726 (second call sees the already-attended audio)
        xa2 = cross_attention(sd, "cross_attn.", xa, xla, xi, xli)
        xi2 = cross_attention(sd, "cross_attn.", xi, xli, xa2, xla)
        return mixer_concat(xi2, xa2, xli, xla)
    raise ValueError(f"Invalid mixer type: {mixer_type}")


# --------------------------------------------------------------------------------------------
# Whole models -- src/transformer/model.py:141-150, 485-543
# --------------------------------------------------------------------------------------------


def unimodal_forward(sd: SD, x, xl, y_in, attn_window: int = -1, dtype=torch.float32):
    """Transformer.forward: model.py:141-150."""
    mem = encode_to_memory(sd, "encoder.", "pos_2d.pe", x.to(dtype))
    return decoder_forward(sd, "decoder.", y_in, mem, xl, attn_window)


def multimodal_memory(sd: SD, xi, xa, xli=None, xla=None, mixer_type="concat", modality="both", dtype=torch.float32):
    """MultimodalTransformer.encoder_forward: model.py:485-522 (modality = the teacher-forcing draw)."""
    mi = encode_to_memory(sd, "image_encoder.", "image_pos_2d.pe", xi.to(dtype))
    ma = encode_to_memory(sd, "audio_encoder.", "audio_pos_2d.pe", xa.to(dtype))
    if modality == "image":
        return mi, xli
    if modality == "audio":
        return ma, xla
    if modality != "both":
        raise ValueError(f"Invalid modality: {modality}")
    return mix(sd, mixer_type, mi, ma, xli, xla)


def multimodal_forward(sd: SD, xi, xli, xa, xla, y_in, mixer_type="concat", attn_window=-1, modality="both", dtype=torch.float32):
    """MultimodalTransformer.forward: model.py:524-543."""
    mem, xl = multimodal_memory(sd, xi, xa, xli, xla, mixer_type, modality, dtype)
    return decoder_forward(sd, "decoder.", y_in, mem, xl, attn_window)


# --------------------------------------------------------------------------------------------
# Greedy decode -- src/transformer/model.py:170-199, 226-262, 592-617
# --------------------------------------------------------------------------------------------


@torch.no_grad()
def greedy_decode(
    sd: SD, memory: torch.Tensor, sos: int, eos: int, max_seq_len: int, attn_window: int = -1, max_steps: Optional[int] = None,
    nhead: int = 4, num_layers: int = 8,
) -> Tuple[List[int], List[float]]:
    """Reference batch-1 loop: re-run the whole decoder on the growing prefix, first-max argmax of
    the last position, append (EOS included) and stop at EOS or after max_seq_len steps.  Returns
    (tokens, raw top logits) -- the latter is what get_pred_seq_and_pred_prob_seq calls "prob"."""
    assert memory.shape[0] == 1, "Inference only supports batch_size = 1"  # model.py:173,238,595
    y_in = torch.tensor([[sos]], dtype=torch.int64)
    toks: List[int] = []
    vals: List[float] = []
    steps = max_seq_len if max_steps is None else min(max_steps, max_seq_len)
    for _ in range(steps):
        logits = decoder_forward(sd, "decoder.", y_in, memory, None, attn_window, nhead, num_layers)[0, :, -1]
        val, tok = logits.max(dim=-1)  # first max index, like argmax / topk(k=1)
        toks.append(int(tok))
        vals.append(float(val))
        if int(tok) == eos:
            break
        y_in = torch.cat([y_in, torch.tensor([[int(tok)]], dtype=torch.int64)], dim=1)
    return toks, vals
