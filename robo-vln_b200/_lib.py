"""ctypes binding of librobovln_b200.so (C ABI: include/robovln_b200.h).

The product path has no fallback: if the shared library is missing and cannot be built the
import fails loudly.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "librobovln_b200.so")


class HcmShape(ctypes.Structure):
    _fields_ = [
        ("B", c_int32), ("N", c_int32), ("L", c_int32), ("instr_rows", c_int32),
        ("rgb_h", c_int32), ("rgb_w", c_int32), ("depth_h", c_int32), ("depth_w", c_int32),
    ]


# name -> (restype, argtypes); the loader checks that every symbol of the header is exported.
SYMBOLS = {
    "hcm_last_error": (c_char_p, []),
    "hcm_version": (c_char_p, []),
    "hcm_dtype": (c_int, []),
    "hcm_create": (c_int, [POINTER(c_void_p)]),
    "hcm_destroy": (None, [c_void_p]),
    "hcm_set_tensor": (c_int, [c_void_p, c_char_p, c_void_p, c_int, c_int, POINTER(c_int64)]),
    "hcm_finalize_weights": (c_int, [c_void_p, c_int, c_int, c_int]),
    "hcm_workspace_bytes": (c_size_t, [c_void_p, POINTER(HcmShape)]),
    "hcm_plan": (c_int, [c_void_p, POINTER(HcmShape), c_void_p, c_size_t]),
    "hcm_forward_hi": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                               c_void_p, c_void_p, c_void_p]),
    "hcm_forward_lo": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                               c_void_p, c_void_p, c_int, c_void_p]),
    "hcm_forward_policy": (c_int, [c_void_p] + [c_void_p] * 5 + [c_int] + [c_void_p] * 9),
    "hcm_forward_policy_host": (c_int, [c_void_p] + [c_void_p] * 12),
    "hcm_set_rgb_format": (c_int, [c_void_p, c_int]),
    "hcm_set_skip_bert": (c_int, [c_void_p, c_int]),
    "hcm_last_launch_count": (c_int64, [c_void_p]),
    "hcm_profile_policy": (c_int, [c_void_p] + [c_void_p] * 5 + [c_int] + [c_void_p] * 7 + [c_char_p, c_size_t, c_void_p]),
    "hcm_run_rgb_trunk": (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    "hcm_run_depth_trunk": (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    "hcm_run_bert": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "hcm_run_encoders": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "hcm_run_cross_modal": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "hcm_get_buffer": (c_int, [c_void_p, c_char_p, POINTER(c_void_p), POINTER(c_int), POINTER(c_int), POINTER(c_int64)]),
    "hcm_copy_buffer": (c_int, [c_void_p, c_char_p, c_void_p, c_size_t, c_void_p]),
    "rvb_conv_gemm": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int64, c_void_p, c_int, c_int, c_int, c_int,
                              c_int, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p, c_int64, c_int, c_int, c_int,
                              c_int, c_int64, c_void_p]),
    "rvb_gemm_ln": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p,
                            c_float, c_void_p, c_int, c_void_p, c_void_p]),
    "rvb_conv_gemm_gn": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                 c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "rvb_rgb_pad_convert": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "rvb_rgb_pad_convert4": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "rvb_groupnorm": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p,
                              c_void_p, c_int64, c_void_p]),
    "rvb_layernorm": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_float, c_void_p, c_int, c_void_p, c_void_p]),
    "rvb_bert_attention": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "rvb_bert_attention_tc": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "rvb_vla_attention": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "rvb_vla_block": (c_int, [c_void_p] * 12 + [c_float, c_int, c_int, c_int, c_void_p, c_int64, c_void_p, c_void_p]),
    "rvb_vla_block_variant": (c_int, [c_int]),
    "rvb_hi_loss": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "rvb_lo_loss": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "rvb_fused_adam": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_double, c_double, c_double,
                               c_double, c_double, c_int, c_double, c_double, c_void_p]),
    "rvb_adam_chunk_elems": (c_int, []),
    "rvb_lstm": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                         c_void_p]),
    "rvb_maxpool3x3s2": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "rvb_rgb_stem_im2col": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "rvb_depth_stem": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "rvb_pack_weight": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_int, c_int,
                                c_int, c_int, c_int64, c_void_p]),
    "rvb_compare_many": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "rvb_checksum": (c_int, [c_void_p, c_size_t, c_void_p, c_void_p]),
}

_libs = {}
LIB_FILES = {"fp16": "librobovln_b200.so", "bf16": "librobovln_b200_bf16.so"}
DTYPE_CODE = {"fp16": 3, "bf16": 1}


def default_dtype() -> str:
    """16-bit operand type of the engine: fp16 unless ROBOVLN_DTYPE=bf16 (csrc/h16.h)."""
    d = os.environ.get("ROBOVLN_DTYPE", "fp16").lower()
    if d in ("fp16", "f16", "half", "float16"):
        return "fp16"
    if d in ("bf16", "bfloat16"):
        return "bf16"
    raise ValueError(f"ROBOVLN_DTYPE={d!r}: expected fp16 or bf16")


def load(build_if_missing: bool = True, dtype: str = None) -> ctypes.CDLL:
    dtype = dtype or default_dtype()
    if dtype in _libs:
        return _libs[dtype]
    path = os.path.join(HERE, LIB_FILES[dtype])
    if not os.path.exists(path):
        if not build_if_missing:
            raise ImportError(f"{path} is missing; run `python robo-vln_b200/build.py`")
        from .build import build  # needs nvcc; raises if it is absent

        build()
    lib = ctypes.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a header symbol
        fn.restype = res
        fn.argtypes = args
    if lib.hcm_dtype() != DTYPE_CODE[dtype]:
        raise ImportError(f"{path} was built for another 16-bit type")
    lib._rvb_dtype = dtype
    _libs[dtype] = lib
    return lib


class HcmError(RuntimeError):
    pass


def check(rc: int, what: str = "", lib=None) -> None:
    if rc != 0:
        # every successful ABI call clears its library's message, so only the failing one is set
        libs = [lib] if lib is not None else (list(_libs.values()) or [load()])
        msg = " | ".join(m.decode() for m in (l.hcm_last_error() for l in libs) if m)
        raise HcmError(f"{what} failed (code {rc}): {msg}")
