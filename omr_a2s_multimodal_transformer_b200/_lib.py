"""ctypes binding of ``libomr_b200.so`` (the C ABI declared in ``include/omr_b200.h``).

There is no fallback: if the shared library is missing or a call fails, a ``RuntimeError`` is raised.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_longlong, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libomr_b200.so")

F32, BF16 = 0, 1

# signature mini-language: i=int, q=long long, p=pointer, f=float, d=double
_SIGS = {
    "omr_cast": "iippqp",
    "omr_relu_bwd": "ipppqp",
    "omr_add": "ipppqp",
    "omr_dropout": "ippqiqfqipp",
    "omr_pack_conv_weight": "ippiiip",
    "omr_pack_dw_weight": "ippip",
    "omr_conv3x3_fwd": "ippppiiiiiiiipp",
    "omr_conv3x3_dgrad": "ipppiiiiiiipfpppp",
    "omr_conv3x3_wgrad": "ippppiiiiiiiipp",
    "omr_dwconv3x3_fwd": "ippppiiiip",
    "omr_dwconv3x3_dgrad": "ipppiiiip",
    "omr_dwconv3x3_wgrad": "ippppiiiiip",
    "omr_instnorm_fwd": "ippppiiifip",
    "omr_instnorm_bwd": "ipppppiiiifipp",
    "omr_pe2d_add": "ipppiiiiiiip",
    "omr_copy_rows": "ippiiiiip",
    "omr_key_bias_from_lengths": "ppiiiifp",
    "omr_key_bias_from_tokens": "ppqqfp",
    "omr_embed_pe_fwd": "ippppiiiipp",
    "omr_embed_bwd": "ipppqiqp",
    "omr_gemm": "iiiiiiipqqpqqpqqipiiip",
    "omr_colsum": "ipqiqpip",
    "omr_attn_next_dropout": "fip",
    "omr_attn_fwd": "ipqqpqqpqqpqqppiiiiifiippip",
    "omr_attn_bwd": "ipqqpqqpqqpqqpqqppqqpqqpqqppiiiiifiippip",
    "omr_add_layernorm_fwd": "ipppppppqifp",
    "omr_layernorm_bwd": "ippppppp" + "qip",
    "omr_dropout_add_layernorm_fwd": "ipppppppqiffqpp",
    "omr_layernorm_bwd_dropout": "ippppppppqifqppp",
    "omr_mask_scale": "ippfqp",
    "omr_ce_fwd": "ipqpqiqppp",
    "omr_ce_reduce": "ppqqpp",
    "omr_ce_bwd": "ipqpppppqiqp",
    "omr_proj_ce_fwd": "ipqpqppqiiqppp",
    "omr_proj_ce_bwd_dx": "ipqpqpppppqiiqpqp",
    "omr_proj_ce_bwd_dw": "ipqpqpppppqiiqppp",
    "omr_adam_tick": "pp",
    "omr_adam_step": "piqpdddddp",
    "omr_argmax_step": "ipqiipppqqppiipp",
    "omr_kv_append": "ipqpiiiipp",
    "omr_attn_decode": "ipqpqqpqqpqpqpqiiiifipp",
    "omr_pad_collate": "pppppiiifpiip",
    "omr_pad_transcripts": "ppiippqp",
    "omr_mix_argmax_step": "ipqpqiifpppqqppiipp",
    "omr_levenshtein": "ppppiippp",
    "omr_decode_persistent": "ipippppiiiiiiiipppppipqqpqfpqpp",
}
_CT = {"i": c_int, "q": c_longlong, "p": c_void_p, "f": c_float, "d": c_double}

_lib = None


def load() -> ctypes.CDLL:
    """Load the CUDA library; raises if it has not been built (no CPU path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -m omr_a2s_multimodal_transformer_b200.build` "
            "(or __graft_entry__.build()); this package has no CPU or PyTorch fallback"
        )
    lib = ctypes.CDLL(LIB_PATH)
    lib.omr_last_error.restype = c_char_p
    lib.omr_abi_version.restype = c_int
    lib.omr_launch_count.restype = c_longlong
    lib.omr_tc_call_count.restype = c_longlong
    lib.omr_decode_persistent_scratch_floats.restype = c_longlong
    lib.omr_decode_persistent_scratch_floats.argtypes = [c_int, c_int, c_int, c_int]
    lib.omr_tensor_core_path_enabled.restype = c_int
    lib.omr_set_tensor_core_path.argtypes = [c_int]
    lib.omr_proj_ce_supported.restype = c_int
    lib.omr_proj_ce_supported.argtypes = [c_int, c_int]
    for name, sig in _SIGS.items():
        fn = getattr(lib, name)
        fn.restype = c_int
        fn.argtypes = [_CT[c] for c in sig]
    _lib = lib
    return lib


def exported_symbols():
    return ["omr_abi_version", "omr_last_error", "omr_launch_count", "omr_tc_call_count", "omr_tensor_core_path_enabled",
            "omr_set_tensor_core_path", "omr_decode_persistent_scratch_floats", "omr_proj_ce_supported"] + list(_SIGS)


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().omr_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed (rc={rc}): {msg}")


def call(name: str, *args) -> None:
    if _prof is None:
        check(getattr(load(), name)(*args), name)
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = launch_count()
    e0.record()
    check(getattr(load(), name)(*args), name)
    e1.record()
    _prof.append((name, _work(name, args), launch_count() - n0, e0, e1, [a for a in args if isinstance(a, int) and abs(a) < (1 << 24)]))


# ---- per-entry-point device timing (bench.py's roofline block; a poor man's timeline) ------------------
_prof = None
_ESZ = {F32: 4, BF16: 2}


def _work(name: str, a) -> tuple:
    """(algorithmic flops, algorithmic bytes) of one C-ABI call, from its arguments."""
    try:
        if name == "omr_gemm":
            m, n, k, batch = a[4], a[5], a[6], a[16]
            return 2.0 * m * n * k * batch, float(batch) * (_ESZ[a[0]] * (m * k + n * k) + _ESZ[a[1]] * m * n)
        if name in ("omr_conv3x3_fwd", "omr_conv3x3_dgrad", "omr_conv3x3_wgrad"):
            if name == "omr_conv3x3_fwd":
                nb, h, w, ci, co, sh, sw = a[5:12]
            elif name == "omr_conv3x3_dgrad":
                nb, h, w, ci, co, sh, sw = a[4:11]
            else:
                nb, h, w, ci, co, sh, sw = a[5:12]
            ho, wo = -(-h // sh), -(-w // sw)
            e = _ESZ[a[0]]
            return 2.0 * nb * ho * wo * co * 9 * ci, float(e) * nb * (h * w * ci + ho * wo * co) + 4.0 * 9 * ci * co
        if name == "omr_attn_fwd":
            b, h, tq, tk, hd = a[15:20]
            return 4.0 * b * h * tq * tk * hd, float(_ESZ[a[0]]) * b * h * hd * (2 * tq + 2 * tk)
        if name == "omr_attn_bwd":
            b, h, tq, tk, hd = a[28:33]
            return 10.0 * b * h * tq * tk * hd, float(_ESZ[a[0]]) * b * h * hd * (4 * tq + 4 * tk)
        if name in ("omr_dwconv3x3_fwd", "omr_dwconv3x3_dgrad"):
            off = 5 if name.endswith("fwd") else 4
            nb, h, w, c = a[off:off + 4]
            return 18.0 * nb * h * w * c, 2.0 * _ESZ[a[0]] * nb * h * w * c
        if name == "omr_dwconv3x3_wgrad":
            nb, h, w, c = a[5:9]
            return 18.0 * nb * h * w * c, 2.0 * _ESZ[a[0]] * nb * h * w * c
        if name == "omr_instnorm_fwd":
            nb, hw, c = a[5:8]
            return 0.0, (2.0 if a[9] else 3.0) * _ESZ[a[0]] * nb * hw * c  # read (stats, unless they came with the conv), read + write (apply)
        if name == "omr_instnorm_bwd":
            nb, hw, c = a[6:9]
            return 0.0, (3.0 if a[11] else 5.0) * _ESZ[a[0]] * nb * hw * c
        if name in ("omr_relu_bwd", "omr_add"):
            return 0.0, 3.0 * _ESZ[a[0]] * a[4]
        if name == "omr_dropout":
            return 0.0, 2.0 * _ESZ[a[0]] * a[3]
        if name == "omr_cast":
            return 0.0, float(_ESZ[a[0]] + _ESZ[a[1]]) * a[4]
        if name == "omr_add_layernorm_fwd":
            return 0.0, 4.0 * _ESZ[a[0]] * a[8] * a[9]
        if name == "omr_layernorm_bwd":
            return 0.0, 3.0 * _ESZ[a[0]] * a[8] * a[9]
        if name == "omr_ce_fwd":
            return 0.0, float(_ESZ[a[0]]) * a[4] * a[5]
        if name == "omr_ce_bwd":
            return 0.0, 2.0 * _ESZ[a[0]] * a[8] * a[9]
        if name == "omr_proj_ce_fwd":  # rows, V, D = a[7:10]: one GEMM; reads x and w, writes two floats per row
            return 2.0 * a[7] * a[8] * a[9], 2.0 * (a[7] + a[8]) * a[9] + 8.0 * a[7]
        if name in ("omr_proj_ce_bwd_dx", "omr_proj_ce_bwd_dw"):  # rows, V, D = a[10:13]: scores recomputed + one gradient GEMM
            return 4.0 * a[10] * a[11] * a[12], 2.0 * (a[10] + a[11]) * a[12] + (2.0 * a[10] if name.endswith("dx") else 4.0 * a[11]) * a[12]
        if name == "omr_adam_step":
            return 0.0, 0.0  # filled in by the caller (needs the parameter count)
        if name == "omr_attn_decode":
            b, h, tk, hd = a[15:19]
            return 4.0 * b * h * tk * hd, 2.0 * _ESZ[a[0]] * b * h * tk * hd
    except Exception:
        pass
    return 0.0, 0.0


def prof_start() -> None:
    """start timing every C-ABI call with CUDA events on the current stream (bench.py only)"""
    global _prof
    _prof = []


def prof_stop(by_shape: bool = False) -> dict:
    """-> {entry point: {calls, launches, ms, flops, bytes}} summed over the calls since prof_start(); with ``by_shape``
    the key is (entry point, tuple of the call's small integer arguments), i.e. one kernel at one launch shape"""
    global _prof
    recs, _prof = _prof or [], None
    torch.cuda.synchronize()
    out = {}
    dump = os.environ.get("OMR_PROF_DUMP")
    if dump:
        with open(dump, "w") as f:
            for name, (fl, by), nl, e0, e1, ints in recs:
                ms = e0.elapsed_time(e1)
                f.write(f"{name}\t{ms:.4f}\t{fl / (ms * 1e-3) / 1e12 if ms > 0 else 0:.1f}\t{by / (ms * 1e-3) / 1e9 if ms > 0 else 0:.0f}\t{ints}\n")
    for name, (fl, by), nl, e0, e1, _ints in recs:
        key = (name, tuple(_ints)) if by_shape else name
        d = out.setdefault(key, {"calls": 0, "launches": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
        d["calls"] += 1
        d["launches"] += nl
        d["ms"] += e0.elapsed_time(e1)
        d["flops"] += fl
        d["bytes"] += by
    return out


def dt_code(dtype: torch.dtype) -> int:
    if dtype == torch.float32:
        return F32
    if dtype == torch.bfloat16:
        return BF16
    raise TypeError(f"unsupported compute dtype {dtype} (float32 or bfloat16)")


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


def launch_count() -> int:
    return int(load().omr_launch_count())


def tc_call_count() -> int:
    return int(load().omr_tc_call_count())


def require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f"{what}: tensor is on {t.device}; omr_a2s_multimodal_transformer_b200 runs only on CUDA (sm_100a) -- "
            "there is no CPU fallback"
        )
