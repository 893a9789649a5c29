"""Batched greedy autoregressive decoding with an in-HBM KV cache.

Replaces the reference's batch-1 loop (``validation_step`` / ``get_pred_seq_and_pred_prob_seq``,
reference model.py:170-199, 226-262, 592-617), which re-runs the whole decoder on the growing prefix
and synchronises with the host at every token.  Here:

* the cross-attention K/V of the encoder memory are projected ONCE per layer and stay resident;
* self-attention K/V rows are appended to a per-layer cache ``[B, Tmax, 2*D]``;
* one decode step is a fixed sequence of kernels that read the position from a device counter, so
  it is captured in a CUDA graph once and replayed; EOS bookkeeping (``finished`` flags, PAD after
  EOS) lives on the device and the host only polls it every ``poll_every`` steps;
* the token stream is identical to the reference loop (first-max argmax, EOS emitted, at most
  ``max_seq_len`` tokens); with ``attn_window > 0`` the cache is read through the same sliding window
  as ``create_variable_window_mask`` (decoder.py:191-217).
"""
from __future__ import annotations

import ctypes
import os
from typing import List, Optional, Tuple

import torch

from . import _lib, ops
from ._lib import ptr, stream_ptr
from .decoder import Decoder
from .params import resolve_dtype


class _DecodeLayer(ctypes.Structure):
    """mirror of omr_decode_layer (include/omr_b200.h)"""

    _fields_ = [(n, ctypes.c_void_p) for n in (
        "w_in", "b_in", "w_o", "b_o", "wc_q", "bc_q", "wc_o", "bc_o", "w1", "b1", "w2", "b2",
        "g1", "be1", "g2", "be2", "g3", "be3", "self_kv", "cross_kv")]


class BatchedGreedyDecoder:
    def __init__(self, decoder: Decoder, dtype: Optional[torch.dtype] = None):
        self.dec = decoder
        self.dtype = resolve_dtype(dtype if dtype is not None else decoder.compute_dtype)
        self._graph = None
        self._graph_key = None

    # -- one decode step (all kernels read the position from the device counter ``pos``) ---------------
    def _step(self, st) -> None:
        logits = self._logits_step(st)
        ops.argmax_step(logits, st["tok"], st["val"], st["finished"], st["eos"], st["pad"], st["out_tokens"],
                        st["out_vals"], 0, step_dev=st["pos"])
        ops.tick(st["pos"])

    def _logits_step(self, st) -> torch.Tensor:
        """the decoder on the current token ``st["tok"]`` at position ``st["pos"]`` -> logits [B,V] (in ``st["logits"]``)"""
        dec, dtype = self.dec, self.dtype
        c = dec._wcache
        b, d, h = st["B"], dec.d_model, dec.nhead
        hd = d // h
        pos = st["pos"]
        esz = st["x"].element_size()
        table = c.get(dec.embedding.weight, "mat", dtype)
        x = ops.embed_pe_fwd(st["tok"].view(b, 1), table, dec.pos_1d.pe.view(-1, d), 0, pos_dev=pos, out=st["x"])
        x2 = x.view(b, d)
        for li, L in enumerate(dec.transformer_decoder.layers):
            sa, ca = L.self_attn, L.multihead_attn
            w_in = c.get(sa.in_proj_weight, "mat", dtype)
            w_o = c.get(sa.out_proj.weight, "mat", dtype)
            wc_in = c.get(ca.in_proj_weight, "mat", dtype)
            wc_o = c.get(ca.out_proj.weight, "mat", dtype)
            w1 = c.get(L.linear1.weight, "mat", dtype)
            w2 = c.get(L.linear2.weight, "mat", dtype)
            qkv = ops.linear_fwd(x2, w_in, sa.in_proj_bias)  # [B,3D]
            cache = st["self_kv"][li]  # [B,Tmax,2D]
            ops.kv_append(qkv.data_ptr() + d * esz, 3 * d, cache, 0, dtype, pos_dev=pos)
            o = torch.empty((b, d), dtype=dtype, device=x.device)
            tmax = cache.shape[1]
            ops.attn_decode(qkv.data_ptr(), 3 * d, cache.data_ptr(), tmax * 2 * d, 2 * d, cache.data_ptr() + d * esz,
                            tmax * 2 * d, 2 * d, o, None, st["ws"], b, h, tmax, hd, dec.attn_window, dtype, pos_dev=pos)
            a = ops.linear_fwd(o, w_o, sa.out_proj.bias)
            x1, _, _ = ops.add_layernorm_fwd(a, x2, L.norm1.weight, L.norm1.bias, L.norm1.eps, False)
            q = ops.linear_fwd(x1, wc_in[:d], ca.in_proj_bias[:d])
            kv = st["cross_kv"][li]  # [B,S,2D]
            s = kv.shape[1]
            o2 = torch.empty((b, d), dtype=dtype, device=x.device)
            ops.attn_decode(q.data_ptr(), d, kv.data_ptr(), s * 2 * d, 2 * d, kv.data_ptr() + d * esz, s * 2 * d, 2 * d, o2,
                            st["mem_bias"], st["ws"], b, h, s, hd, 0, dtype)
            cc = ops.linear_fwd(o2, wc_o, ca.out_proj.bias)
            x2n, _, _ = ops.add_layernorm_fwd(cc, x1, L.norm2.weight, L.norm2.bias, L.norm2.eps, False)
            hmid = ops.linear_fwd(x2n, w1, L.linear1.bias, relu=True)
            f = ops.linear_fwd(hmid, w2, L.linear2.bias)
            x2, _, _ = ops.add_layernorm_fwd(f, x2n, L.norm3.weight, L.norm3.bias, L.norm3.eps, False)
        wout = c.get(dec.out_layer.weight, "mat", dtype)
        return ops.linear_fwd(x2, wout, dec.out_layer.bias, out=st["logits"])

    def _persistent_ok(self, b: int) -> bool:
        dec = self.dec
        return (os.environ.get("OMR_DECODE_MODE", "persistent") != "graph" and dec.d_model == 256 and dec.nhead == 4
                and dec.ff_dim == 256)

    def _decode_persistent(self, st, cross_kv, steps: int, stop_at_eos: bool, poll_every: int) -> int:
        """the whole step loop as ONE launch of the persistent kernel (csrc/decode_persistent.cu): a 4-CTA cluster per
        sample runs all steps of that sample and stops at its own EOS -- no host polling.  Returns the step budget."""
        dec, dtype = self.dec, self.dtype
        c = dec._wcache
        b, d = st["B"], dec.d_model
        dev = st["tok"].device
        keep = []  # keeps sliced weight views alive until the launches are enqueued
        recs = (_DecodeLayer * len(cross_kv))()
        # the persistent kernel streams ONE head per CTA: its caches are head-major, [B][H][2][rows][hd] (the K rows of a head
        # are one contiguous block, its V rows the next); the per-kernel path keeps the [B, rows, 2D] caches of the state
        h, hd = dec.nhead, d // dec.nhead
        s_mem, t_max = cross_kv[0].shape[1], st["self_kv"][0].shape[1]
        cross_hm = [kv.view(b, s_mem, 2, h, hd).permute(0, 3, 2, 1, 4).contiguous() for kv in cross_kv]
        self_hm = [torch.zeros((b, h, 2, t_max, hd), dtype=dtype, device=dev) for _ in cross_kv]
        for li, L in enumerate(dec.transformer_decoder.layers):
            sa, ca = L.self_attn, L.multihead_attn
            # layouts of the persistent kernel (include/omr_b200.h, omr_decode_layer): bf16 = mma.sync A-fragment order,
            # fp32 = row-major; the projections that FOLLOW an attention head / the FFN quarter are split along their
            # reduction index inside the kernel and come as four column slices
            ns, ks = ("matDecA", "matDecKS") if dtype == torch.bfloat16 else ("mat", "matKS4")
            wc_in = c.get(ca.in_proj_weight, ns, dtype)
            vals = dict(
                w_in=c.get(sa.in_proj_weight, ns, dtype), b_in=sa.in_proj_bias,
                w_o=c.get(sa.out_proj.weight, ks, dtype), b_o=sa.out_proj.bias,
                wc_q=wc_in[:d], bc_q=ca.in_proj_bias[:d],
                wc_o=c.get(ca.out_proj.weight, ks, dtype), bc_o=ca.out_proj.bias,
                w1=c.get(L.linear1.weight, ns, dtype), b1=L.linear1.bias,
                w2=c.get(L.linear2.weight, ks, dtype), b2=L.linear2.bias,
                g1=L.norm1.weight, be1=L.norm1.bias, g2=L.norm2.weight, be2=L.norm2.bias, g3=L.norm3.weight, be3=L.norm3.bias,
                self_kv=self_hm[li], cross_kv=cross_hm[li])
            for k, t in vals.items():
                setattr(recs[li], k, t.data_ptr())
                keep.append(t)
        table = torch.frombuffer(bytearray(bytes(recs)), dtype=torch.uint8).to(dev)
        nfl = int(_lib.load().omr_decode_persistent_scratch_floats(b, dec.nhead, d, dec.output_size))
        scratch = torch.empty(nfl, dtype=torch.float32, device=dev)
        wout = c.get(dec.out_layer.weight, "matDecA" if dtype == torch.bfloat16 else "mat", dtype)  # bf16: rows padded to 32
        table_emb = c.get(dec.embedding.weight, "mat", dtype)
        mem_bias = st["mem_bias"]
        s_len = cross_kv[0].shape[1]
        tmax = st["self_kv"][0].shape[1]
        timing = torch.zeros(32, dtype=torch.int64, device=dev) if os.environ.get("OMR_DECODE_TIMING") else None
        _lib.call("omr_decode_persistent", _lib.dt_code(dtype), ptr(table), len(cross_kv), ptr(table_emb),
                  ptr(dec.pos_1d.pe), ptr(wout), ptr(dec.out_layer.bias), b, dec.nhead, d, dec.output_size, s_len, tmax, steps,
                  dec.attn_window if dec.attn_window and dec.attn_window > 0 else 0, ptr(st["tok"]), ptr(st["val"]),
                  ptr(st["finished"]), ptr(st["out_tokens"]), ptr(st["out_vals"]), st["out_tokens"].shape[1], ptr(st["pos"]),
                  st["eos"], st["pad"], ptr(mem_bias), mem_bias.stride(0) if mem_bias is not None else 0,
                  float(dec.transformer_decoder.layers[0].norm1.eps), ptr(scratch), nfl, ptr(timing), stream_ptr())
        done = steps
        if stop_at_eos:  # trim the all-PAD tail (every sample stopped at its own EOS inside the kernel)
            nz = (st["out_tokens"] != st["pad"]).any(dim=0).nonzero()
            done = int(nz.max().item()) + 1 if nz.numel() else 1
        del keep
        if timing is not None:
            names = ["embed", "qkv", "self_attn", "out_proj", "cross_q", "cross_attn", "cross_out", "ffn1", "ffn2", "classifier", "argmax"]
            cyc = timing.cpu().tolist()
            print(f"[decode timing] of which cluster-barrier wait {cyc[11] / max(done, 1):.0f}, weight-ring wait {cyc[12] / max(done, 1):.0f}, ring top-up at barriers {cyc[13] / max(done, 1):.0f}, barrier arrive (release) {cyc[14] / max(done, 1):.0f} cycles per step")
            ph = {"0": "cross-q", "1": "q|k|v", "2": "FFN1"}.get(os.environ.get("OMR_DECODE_DBG_PHASE", "0"), "cross-q")
            print(f"[decode timing] {ph} projection, cycles per step (input vector by warp 0 + barrier, -, MMAs + partial sums, "
                  "barrier, epilogue + slot release):", [round(c / max(done, 1)) for c in cyc[16:21]])
            cyc = cyc[:11]
            tot = sum(cyc) or 1
            print("[decode timing, SM cycles per step on CTA 0] " + ", ".join(
                f"{n} {c / max(done, 1):.0f} ({100 * c / tot:.0f}%)" for n, c in zip(names, cyc)) + f" | total {tot / max(done, 1):.0f}")
        return done

    def _new_state(self, memory: torch.Tensor, sos: int, eos: int, pad: int, max_steps: Optional[int], stop_at_eos: bool,
                   mem_bias: Optional[torch.Tensor], shared: Optional[dict] = None) -> dict:
        """cross K/V of ``memory`` (projected once per layer), the self K/V cache and the device-side bookkeeping of one
        decode; ``shared`` = a state whose token / position / output buffers this one reuses (lock-step late fusion)"""
        dec, dtype = self.dec, self.dtype
        ops._lib.require_cuda(memory, "greedy decode")
        dev = memory.device
        mem = ops.cast(memory.contiguous(), dtype)
        b, s, d = mem.shape
        steps = dec.max_seq_len if max_steps is None else min(int(max_steps), dec.max_seq_len)
        if shared is not None:
            steps = shared["steps"]
        c = dec._wcache
        mem2 = mem.view(b * s, d)
        cross_kv = []
        for L in dec.transformer_decoder.layers:
            ca = L.multihead_attn
            wc_in = c.get(ca.in_proj_weight, "mat", dtype)
            cross_kv.append(ops.linear_fwd(mem2, wc_in[d:], ca.in_proj_bias[d:]).view(b, s, 2 * d))
        nl = len(cross_kv)
        st = {
            "B": b, "steps": steps, "eos": int(eos), "pad": int(pad), "cross_kv": cross_kv, "mem_bias": mem_bias,
            "self_kv": [torch.zeros((b, steps, 2 * d), dtype=dtype, device=dev) for _ in range(nl)],
            "x": torch.empty((b, 1, d), dtype=dtype, device=dev),
            "logits": torch.empty((b, dec.output_size), dtype=dtype, device=dev),
            "ws": torch.empty(ops.attn_decode_ws_floats(b, dec.nhead), dtype=torch.float32, device=dev),
        }
        if shared is not None:
            for k in ("tok", "val", "finished", "out_tokens", "out_vals", "pos"):
                st[k] = shared[k]
        else:
            st.update({
                "tok": torch.full((b,), int(sos), dtype=torch.int64, device=dev),
                "val": torch.zeros((b,), dtype=torch.float32, device=dev),
                "finished": torch.zeros((b,), dtype=torch.int32, device=dev),
                "out_tokens": torch.full((b, steps), int(pad), dtype=torch.int64, device=dev),
                "out_vals": torch.zeros((b, steps), dtype=torch.float32, device=dev),
                "pos": torch.zeros((1,), dtype=torch.int32, device=dev),
            })
        if not stop_at_eos:
            st["eos"] = -1  # never matches: forced full-length decoding
        return st

    @torch.no_grad()
    def decode(self, memory: torch.Tensor, sos: int, eos: int, pad: int = 0, max_steps: Optional[int] = None,
               stop_at_eos: bool = True, use_graph: bool = True, poll_every: int = 64,
               mem_bias: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """memory [B,S,D] -> (tokens int64 [B,steps] (PAD after EOS), top logits fp32 [B,steps], lengths int64 [B])."""
        dec = self.dec
        st = self._new_state(memory, sos, eos, pad, max_steps, stop_at_eos, mem_bias)
        b, steps, dev, cross_kv = st["B"], st["steps"], memory.device, st["cross_kv"]

        def reset() -> None:
            st["tok"].fill_(int(sos))
            st["finished"].zero_()
            st["pos"].zero_()
            st["out_tokens"].fill_(int(pad))
            st["out_vals"].zero_()

        if use_graph and self._persistent_ok(b):
            done_steps = self._decode_persistent(st, cross_kv, steps, stop_at_eos, poll_every)
            toks = st["out_tokens"][:, :done_steps]
            vals = st["out_vals"][:, :done_steps]
            is_eos = toks == int(eos) if stop_at_eos else torch.zeros_like(toks, dtype=torch.bool)
            first = torch.where(is_eos.any(dim=1), is_eos.float().argmax(dim=1) + 1, torch.full((b,), done_steps, device=dev))
            return toks, vals, first.to(torch.int64)
        return _drive(lambda: self._step(st), st, reset, steps, int(eos), stop_at_eos, use_graph, poll_every, dev)

    @staticmethod
    def to_lists(tokens: torch.Tensor, vals: torch.Tensor, lengths: torch.Tensor) -> Tuple[List[List[int]], List[List[float]]]:
        tk, vl, ln = tokens.cpu().tolist(), vals.cpu().tolist(), lengths.cpu().tolist()
        return [r[:n] for r, n in zip(tk, ln)], [r[:n] for r, n in zip(vl, ln)]


def _drive(step_fn, st, reset, steps: int, eos: int, stop_at_eos: bool, use_graph: bool, poll_every: int, dev):
    """run ``step_fn`` (one decode step reading the device-side position) ``steps`` times -- replayed from a CUDA graph
    captured once -- polling the ``finished`` flags every ``poll_every`` steps; -> (tokens, values, lengths)"""
    b = st["B"]
    graph = None
    if use_graph and steps > 2:
        # warm-up on a side stream (sets kernel attributes, fills the weight cache), then capture one step
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            step_fn()
        torch.cuda.current_stream(dev).wait_stream(side)
        reset()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            step_fn()
        reset()
    done_steps = 0
    for t in range(steps):
        if graph is not None:
            graph.replay()
        else:
            step_fn()
        done_steps = t + 1
        if stop_at_eos and poll_every > 0 and done_steps % poll_every == 0 and done_steps < steps:
            if bool(st["finished"].all().item()):
                break
    toks = st["out_tokens"][:, :done_steps]
    vals = st["out_vals"][:, :done_steps]
    is_eos = toks == int(eos) if stop_at_eos else torch.zeros_like(toks, dtype=torch.bool)
    first = torch.where(is_eos.any(dim=1), is_eos.float().argmax(dim=1) + 1, torch.full((b,), done_steps, device=dev))
    return toks, vals, first.to(torch.int64)


class WeightedGreedyDecoder:
    """Token-level late fusion of two unimodal decoders (reference ``weighted_prediction``,
    src/multimodal/weighted_multimodal/test.py:21-70): both KV-cached decoders are stepped in lock-step on the SAME
    token; each step mixes their vocabulary distributions ``alpha * softmax(img) + (1 - alpha) * softmax(audio)`` and
    takes the first-max argmax (``omr_mix_argmax_step``).  Batched; the reference handles one sample at a time."""

    def __init__(self, img_decoder: Decoder, audio_decoder: Decoder, dtype: Optional[torch.dtype] = None):
        if img_decoder.output_size != audio_decoder.output_size:
            raise ValueError("Vocabularies do not match")  # the reference asserts w2i equality (test.py:140)
        self.a = BatchedGreedyDecoder(img_decoder, dtype)
        self.b = BatchedGreedyDecoder(audio_decoder, dtype)
        self.dtype = self.a.dtype

    @torch.no_grad()
    def decode(self, memory_img: torch.Tensor, memory_audio: torch.Tensor, sos: int, eos: int, pad: int = 0,
               alpha: float = 0.5, max_steps: Optional[int] = None, stop_at_eos: bool = True, use_graph: bool = True,
               poll_every: int = 64) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """memories [B,S_i,D], [B,S_a,D] -> (tokens int64 [B,steps], mixed probabilities fp32 [B,steps], lengths [B])"""
        if memory_img.shape[0] != memory_audio.shape[0]:
            raise ValueError("image and audio batches differ")
        # the reference loops max(img.max_seq_len, audio.max_seq_len) times; both PE tables must cover the prefix
        limit = min(self.a.dec.max_seq_len, self.b.dec.max_seq_len)
        steps = limit if max_steps is None else min(int(max_steps), limit)
        sa = self.a._new_state(memory_img, sos, eos, pad, steps, stop_at_eos, None)
        sb = self.b._new_state(memory_audio, sos, eos, pad, steps, stop_at_eos, None, shared=sa)
        dev = memory_img.device

        def step() -> None:
            la = self.a._logits_step(sa)
            lb = self.b._logits_step(sb)
            ops.mix_argmax_step(la, lb, alpha, sa["tok"], sa["val"], sa["finished"], sa["eos"], sa["pad"], sa["out_tokens"],
                                sa["out_vals"], 0, step_dev=sa["pos"])
            ops.tick(sa["pos"])

        def reset() -> None:
            sa["tok"].fill_(int(sos))
            sa["finished"].zero_()
            sa["pos"].zero_()
            sa["out_tokens"].fill_(int(pad))
            sa["out_vals"].zero_()

        return _drive(step, sa, reset, sa["steps"], int(eos), stop_at_eos, use_graph, poll_every, dev)
