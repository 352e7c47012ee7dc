"""ctypes binding of ``libgenvc_b200.so`` (C ABI: ``include/genvc_b200.h``).

The product path has no CPU fallback: if the shared library is missing or a call
fails, a ``RuntimeError`` is raised.  Build it with ``python -m genvc_b200.build``
(or ``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libgenvc_b200.so")

GENVC_OK = 0
GENVC_E_INVALID = -1
GENVC_E_STATE = -2
GENVC_E_CUDA = -3
GENVC_E_UNSUPPORTED = -4


class GenvcConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "n_layer", "d_model", "n_head", "n_text_vocab", "n_audio_vocab",
        "start_text", "stop_text", "start_audio", "stop_audio",
        "n_mel_pos", "n_text_pos", "max_gen_mel_tokens",
        "pc_depth", "pc_dim_context", "pc_latents", "pc_dim_head", "pc_heads", "pc_ff_inner",
        "max_batch", "max_seq", "max_mel_frames",
    )]


class GenvcSampling(C.Structure):
    _fields_ = [
        ("top_k", C.c_int32),
        ("top_p", C.c_float),
        ("top_p_threshold", C.c_float),
        ("temperature", C.c_float),
        ("repetition_penalty", C.c_float),
        ("ignore_eos", C.c_int32),
        ("max_new_tokens", C.c_int32),
        ("seed", C.c_uint64),
    ]


# name -> (restype, argtypes); every symbol include/genvc_b200.h declares
_P = C.c_void_p
SIGNATURES = {
    "genvc_create": (C.c_int, [C.POINTER(GenvcConfig), C.c_int, C.POINTER(_P)]),
    "genvc_destroy": (None, [_P]),
    "genvc_last_error": (C.c_char_p, [_P]),
    "genvc_decode_grid": (C.c_int, [_P]),
    "genvc_fused_max_rows": (C.c_int, [_P]),
    "genvc_blob_floats": (C.c_uint64, [_P]),
    "genvc_num_tensors": (C.c_int, [_P]),
    "genvc_tensor_name": (C.c_int, [_P, C.c_int, C.c_char_p, C.c_size_t]),
    "genvc_tensor_info": (C.c_int, [_P, C.c_char_p] + [C.POINTER(C.c_uint64)] * 4),
    "genvc_bind_weights": (C.c_int, [_P, _P, C.c_uint64]),
    "genvc_stream_floats": (C.c_uint64, [_P]),
    "genvc_pack_stream": (C.c_int, [_P, _P, C.c_uint64, _P]),
    "genvc_tc_floats": (C.c_uint64, [_P]),
    "genvc_pack_tc": (C.c_int, [_P, _P, C.c_uint64, _P]),
    "genvc_kv_floats": (C.c_uint64, [_P]),
    "genvc_workspace_bytes": (C.c_uint64, [_P]),
    "genvc_bind_buffers": (C.c_int, [_P, _P, C.c_uint64, _P, C.c_uint64]),
    "genvc_vw_floats": (C.c_uint64, [_P]),
    "genvc_bind_vw": (C.c_int, [_P, _P, C.c_uint64]),
    "genvc_perceiver": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, _P]),
    "genvc_embed_prefix": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, _P, _P]),
    "genvc_prefill": (C.c_int, [_P, _P, C.c_int, C.c_int, _P]),
    "genvc_decode": (C.c_int, [_P, C.c_int, C.POINTER(GenvcSampling), _P, _P, _P, _P, _P, _P, C.c_int, _P]),
    "genvc_forward_latents": (C.c_int, [_P, _P, _P, C.c_int, _P, C.c_int, C.c_int, _P, _P]),
    "genvc_kv_attention": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P]),
    "genvc_launch_count": (C.c_uint64, [_P]),
    "genvc_conv1d": (C.c_int, [_P, _P, _P, _P, _P] + [C.c_int] * 8 + [C.c_float, C.c_int, C.c_float, C.c_int, _P, C.c_uint64, _P]),
    "genvc_codebook_argmin": (C.c_int, [_P, _P, _P] + [C.c_int] * 4 + [_P]),
    "genvc_mel_spectrogram": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, _P, _P, _P] + [C.c_int] * 5 + [C.c_float, _P]),
    "genvc_conv_transpose1d": (C.c_int, [_P, _P, _P, _P] + [C.c_int] * 7 + [C.c_float, _P, C.c_uint64, _P]),
    "genvc_debug_layout": (C.c_int, [_P, _P, C.c_int]),
    "genvc_debug_trace": (C.c_int, [_P, _P, C.c_int, C.c_int]),
    "genvc_debug_tune": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
}

_lib = None


def load_library(path: str | None = None) -> C.CDLL:
    """dlopen the library and type every entry point.  Raises if it is absent."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("GENVC_B200_LIB", LIB_PATH)
    if not os.path.exists(p):
        raise RuntimeError(
            f"{p} not found: the CUDA library is not built (run `python -m genvc_b200.build`). "
            "genvc_b200 has no CPU fallback."
        )
    lib = C.CDLL(p)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


class GenvcError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"genvc_b200 error {code}: {msg}")
        self.code = code
