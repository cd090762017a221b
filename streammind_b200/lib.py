"""ctypes binding of libstreammind_b200.so (the C ABI declared in include/streammind_b200.h).

The CUDA library is the ONLY implementation of the hot path: if it is missing or cannot be loaded
this module raises -- there is no CPU or PyTorch fallback."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SMB_LIB_PATH") or os.path.join(_HERE, "libstreammind_b200.so")   # override: A/B builds

SM_DTYPE_F16, SM_DTYPE_BF16, SM_DTYPE_F32 = 0, 1, 2


class SmConfig(C.Structure):
    """Mirror of ``struct sm_config`` (include/streammind_b200.h) -- keep field order in sync."""
    _fields_ = [
        ("dtype", C.c_int), ("max_frames", C.c_int),
        ("vit_image", C.c_int), ("vit_patch", C.c_int), ("vit_hidden", C.c_int), ("vit_layers", C.c_int),
        ("vit_heads", C.c_int), ("vit_ffn", C.c_int), ("vit_eps", C.c_float),
        ("proj_d_model", C.c_int), ("proj_d_state", C.c_int), ("proj_d_conv", C.c_int), ("proj_expand", C.c_int),
        ("proj_eps", C.c_float),
        ("gate_layers", C.c_int), ("gate_heads", C.c_int), ("gate_kv_heads", C.c_int), ("gate_head_dim", C.c_int),
        ("gate_ffn", C.c_int), ("gate_eps", C.c_float),
        ("llm_hidden", C.c_int), ("llm_layers", C.c_int), ("llm_heads", C.c_int), ("llm_kv_heads", C.c_int),
        ("llm_head_dim", C.c_int), ("llm_ffn", C.c_int), ("llm_vocab", C.c_int), ("llm_max_ctx", C.c_int),
        ("llm_eps", C.c_float), ("llm_rope_theta", C.c_float),
        ("use_graphs", C.c_int), ("n_streams", C.c_int),
    ]


# symbol -> (restype, argtypes); every symbol include/streammind_b200.h declares
_VP, _I, _LL = C.c_void_p, C.c_int, C.c_longlong
SYMBOLS = {
    "sm_create": (_I, [C.POINTER(_VP), _I, C.POINTER(SmConfig)]),
    "sm_destroy": (None, [_VP]),
    "sm_last_error": (C.c_char_p, [_VP]),
    "sm_load_weight": (_I, [_VP, C.c_char_p, _VP, _I, _I, _I, C.POINTER(C.c_int64)]),
    "sm_finalize_weights": (_I, [_VP]),
    "sm_stream_reset": (_I, [_VP, _VP]),
    "sm_stream_select": (_I, [_VP, _I]),
    "sm_num_streams": (_I, [_VP]),
    "sm_vit_encode": (_I, [_VP, _VP, _I, _VP, _VP, _VP]),
    "sm_pool_features": (_I, [_VP, _VP, _I, _VP, _VP]),
    "sm_projector_step": (_I, [_VP, _VP, _I, _VP, _VP]),
    "sm_gate_score": (_I, [_VP, _VP, _VP, _VP]),
    "sm_frame_step": (_I, [_VP, _VP, _I, _I, _VP, _VP, _VP, _VP, _VP]),
    "sm_frame_step_multi": (_I, [_VP, _VP, _I, _I, _I, _VP, _VP, _VP, _VP]),
    "sm_frame_submit": (_I, [_VP, _VP, _I, _I, _VP, _VP, _VP, _VP, _VP, C.POINTER(C.c_longlong)]),
    "sm_frame_wait": (_I, [_VP, _LL, _VP, _I]),
    "sm_cognition_sample": (_I, [_VP, _VP, _I, _I, _I, C.c_double, _VP, _VP, _VP]),
    "sm_cognition_count": (_I, [_I, C.c_double, _I]),
    "sm_linspace_indices": (_I, [_I, _I, _I, _VP]),
    "sm_embed_tokens": (_I, [_VP, _VP, _I, _VP, _VP]),
    "sm_llm_prefill": (_I, [_VP, _VP, _I, _VP, _VP]),
    "sm_llm_decode": (_I, [_VP, _I, _VP, _I, _VP, _VP, _VP]),
    "sm_llm_decode_multi": (_I, [_VP, _I, _VP, _VP, _VP, _I, _VP, _I, _VP, _VP]),
    "sm_decode_stats": (_I, [_VP, C.POINTER(C.c_double), C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), _I]),
    "sm_debug_decode_logits": (_I, [_VP, _I, _VP, _VP]),
    "sm_debug_decode_phases": (_I, [_VP, _VP]),
    "sm_kv_len": (_I, [_VP]),
    "sm_kv_set_len": (_I, [_VP, _I]),
    "sm_test_gemm": (_I, [_VP, _VP, _VP, _VP, _VP, _I, _I, _I, _I, _I, _I, _VP]),
    "sm_test_gemm_trace": (_I, [_VP, _VP]),
    "sm_test_attention": (_I, [_VP, _VP, _VP, _I, _I, _I, _I, _VP]),
    "sm_test_kv_attention": (_I, [_VP, _VP, _I, _VP, _VP, _I, _VP, _I, _I, _I, _I, _I, _VP]),
    "sm_profile_enable": (_I, [_VP, _I]),
    "sm_profile_read": (_I, [_VP, _I, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]),
    "sm_profile_class_name": (C.c_char_p, [_I]),
    "sm_debug_kernel_filter": (_I, [_VP, C.c_uint]),
    "sm_debug_attention_mode": (_I, [_VP, _I]),
    "sm_resample_table": (_I, [_I, _I, C.POINTER(C.c_int), _VP, _VP, _LL]),
    "sm_preprocess_frames": (_I, [_VP, _VP, _I, _I, _I, _I, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_int), _VP, _VP]),
    "sm_launch_count": (_LL, [_VP, _I]),
}

_lib = None


def load() -> C.CDLL:
    """Load the shared library and bind every declared symbol; raises if anything is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(streammind_b200 has no CPU / PyTorch fallback for the hot path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
