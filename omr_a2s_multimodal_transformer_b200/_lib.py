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
    "omr_dropout": "ippqiqfqip",
    "omr_pack_conv_weight": "ippiiip",
    "omr_pack_dw_weight": "ippip",
    "omr_conv3x3_fwd": "ippppiiiiiiiip",
    "omr_conv3x3_dgrad": "ipppiiiiiiip",
    "omr_conv3x3_wgrad": "ippppiiiiiiiip",
    "omr_dwconv3x3_fwd": "ippppiiiip",
    "omr_dwconv3x3_dgrad": "ipppiiiip",
    "omr_dwconv3x3_wgrad": "ippppiiiiip",
    "omr_instnorm_fwd": "ipppiiifp",
    "omr_instnorm_bwd": "ipppppiiip",
    "omr_pe2d_add": "ipppiiiiiiip",
    "omr_copy_rows": "ippiiiiip",
    "omr_key_bias_from_lengths": "ppiiiifp",
    "omr_key_bias_from_tokens": "ppqqfp",
    "omr_embed_pe_fwd": "ippppiiiipp",
    "omr_embed_bwd": "ipppqiqp",
    "omr_gemm": "iiiiiiipqqpqqpqqipiiip",
    "omr_colsum": "ipqiqpip",
    "omr_attn_fwd": "ipqqpqqpqqpqqppiiiiifiippip",
    "omr_attn_bwd": "ipqqpqqpqqpqqpqqppqqpqqpqqppiiiiifiippip",
    "omr_add_layernorm_fwd": "ipppppppqifp",
    "omr_layernorm_bwd": "ippppppp" + "qip",
    "omr_ce_fwd": "ipqpqiqppp",
    "omr_ce_reduce": "ppqqpp",
    "omr_ce_bwd": "ipqpppppqiqp",
    "omr_adam_tick": "pp",
    "omr_adam_step": "piqpdddddp",
    "omr_argmax_step": "ipqiipppqqppiipp",
    "omr_kv_append": "ipqpiiiipp",
    "omr_attn_decode": "ipqpqqpqqpqpqpqiiiifipp",
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
    lib.omr_tensor_core_path_enabled.restype = c_int
    lib.omr_set_tensor_core_path.argtypes = [c_int]
    for name, sig in _SIGS.items():
        fn = getattr(lib, name)
        fn.restype = c_int
        fn.argtypes = [_CT[c] for c in sig]
    _lib = lib
    return lib


def exported_symbols():
    return ["omr_abi_version", "omr_last_error", "omr_launch_count", "omr_tensor_core_path_enabled"] + list(_SIGS)


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().omr_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed (rc={rc}): {msg}")


def call(name: str, *args) -> None:
    check(getattr(load(), name)(*args), name)


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


def require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f"{what}: tensor is on {t.device}; omr_a2s_multimodal_transformer_b200 runs only on CUDA (sm_100a) -- "
            "there is no CPU fallback"
        )
