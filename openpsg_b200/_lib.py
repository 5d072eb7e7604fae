"""ctypes loader for libopsg_b200.so (the C-ABI library declared in include/opsg_b200.h).

There is deliberately no fallback: if the library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_void_p
from pathlib import Path

_LIB_PATH = Path(__file__).resolve().parent / "libopsg_b200.so"
_lib = None

OPSG_OK, OPSG_E_INVALID, OPSG_E_CUDA, OPSG_E_NO_DEVICE, OPSG_E_UNSUPPORTED = 0, -1, -2, -3, -4
ACT_NONE, ACT_GELU, ACT_RELU = 0, 1, 2
OUT_BF16, OUT_F32, OUT_F32_ATOMIC = 0, 1, 2

P, I, F = c_void_p, c_int, c_float

# name -> argtypes, exactly the prototypes of include/opsg_b200.h (tests/test_cabi.py cross-checks the header)
SIGNATURES = {
    "opsg_version": [],
    "opsg_last_error_string": [],
    "opsg_device_check": [],
    "opsg_num_sms": [],
    "opsg_pair_mask_bits": [P, I, I, I, I, I, I, I, I, P, I, P, I, P],
    "opsg_patch_im2col": [P, I, I, I, I, P, P],
    "opsg_gemm_bf16": [P, I, P, I, P, I, I, I, I, P, I, P, I, I, I, I, P],
    "opsg_gemm_bf16_ln": [P, I, P, I, P, I, I, I, I, P, P, I, I, P, P, P, P, P, P, F, P],
    "opsg_gemm_streamk_workspace_bytes": [I, I],
    "opsg_gemm_bf16_streamk": [P, I, P, I, P, I, I, I, I, P, P, I, I, I, P, ctypes.c_size_t, P],
    "opsg_cast_f32_bf16": [P, I, P, I, I, I, P],
    "opsg_init_rows_f32": [P, I, P, I, I, P],
    "opsg_qformer_embed_ln": [P, I, P, I, I, P, I, P, P, P, F, I, P, P],
    "opsg_layernorm_bf16": [P, P, P, F, P, I, I, P],
    "opsg_self_attn_small": [P, P, P, I, I, I, I, I, I, P, P],
    "opsg_token_order": [P, I, I, I, P, P, P],
    "opsg_xattn_bias_tiles_bytes": [I, I],
    "opsg_xattn_bias_tiles": [P, I, P, I, I, I, I, P, P],
    "opsg_xattn_pairs": [P, P, I, P, I, P, I, P, I, I, I, I, I, I, P, P, P],
    "opsg_exist_filter_topk": [P, I, I, I, P, P, F, I, P, P, P, P, P],
    "opsg_mask_pool_labels": [P, I, I, I, I, I, I, I, I, P, I, P, P, P],
    "opsg_mask_pool_workspace_bytes": [I, I, I, I],
    "opsg_mask_pool_pairs": [P, I, I, I, P, P, I, P, P, I, I, I, P, ctypes.c_size_t, P, P, P],
    "opsg_gather_rows_bf16": [P, I, P, I, P, P],
    "opsg_embed_gather": [P, I, P, P, P, I, P, I, P],
    "opsg_llm_build_prefix": [P, I, I, I, P, P, I, P, P, I, I, P, P],
    "opsg_llm_attn": [P, I, P, P, I, P, I, I, I, I, I, F, P, I, P],
    "opsg_llm_attn_append": [P, I, P, P, I, P, I, I, I, I, F, P, I, P],
    "opsg_kv_append": [P, I, I, I, I, I, P, P, I, P],
    "opsg_argmax_rows": [P, I, I, I, P, P],
    "opsg_rmsnorm_bf16": [P, I, P, F, P, I, I, I, P],
    "opsg_rope_bf16": [P, I, I, I, I, I, P, P, P, I, P],
    "opsg_swiglu_bf16": [P, I, I, I, P, I, P],
    "opsg_llm_prompt_layout": [P, I, I, I, I, I, P, P, P, P, P],
    "opsg_copy_bytes": [P, P, ctypes.c_size_t, P],
    "opsg_transpose_i32": [P, I, I, P, P],
    "opsg_splitk_reduce_bf16": [P, I, I, I, P, P, I, P, I, P],
    "opsg_pan_relabel": [P, ctypes.c_longlong, P, P, I, P, P],
    "opsg_pan_colorize": [P, ctypes.c_longlong, P, P, I, P, P],
}


class OpsgError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libopsg_b200 error {code}: {message}")
        self.code = code


def library_path() -> Path:
    return Path(os.environ.get("OPSG_B200_LIB", str(_LIB_PATH)))


def load():
    """Load the shared library (once).  Raises if it has not been built: run ``__graft_entry__.build()``
    or ``make -C openpsg_b200/csrc``."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not path.exists():
        raise OpsgError(OPSG_E_INVALID, f"{path} not found: build it with `make -C openpsg_b200/csrc` "
                                       "(there is no non-CUDA fallback)")
    lib = ctypes.CDLL(str(path))
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = (c_char_p if name == "opsg_last_error_string"
                      else ctypes.c_size_t if name in ("opsg_xattn_bias_tiles_bytes", "opsg_gemm_streamk_workspace_bytes",
                                                       "opsg_mask_pool_workspace_bytes")
                      else c_int)
    _lib = lib
    return lib


def last_error() -> str:
    return load().opsg_last_error_string().decode("utf-8", "replace")


def check(rc: int) -> None:
    if rc != OPSG_OK:
        raise OpsgError(rc, last_error())
