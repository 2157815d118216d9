"""ctypes binding of the C ABI declared in include/kbner_b200.h.

There is NO fallback: if libkbner_b200.so is missing or a call fails, this raises.  The CPU
oracle under oracle/ is test infrastructure and is never imported from here.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libkbner_b200.so")

_c_void_p, _c_int, _c_float = ctypes.c_void_p, ctypes.c_int, ctypes.c_float

# name -> argtypes ; every entry point of include/kbner_b200.h
SIGNATURES = {
    "kbner_abi_version": ([], _c_int),
    "kbner_last_error": ([], ctypes.c_char_p),
    "kbner_device_check": ([_c_int], _c_int),
    "kbner_launch_count": ([], ctypes.c_uint64),
    "kbner_add_launches": ([ctypes.c_uint64], None),
    "kbner_set_sm_budget": ([_c_int], _c_int),
    "kbner_crf_compact": ([_c_void_p, _c_int, _c_int, _c_void_p, _c_void_p, _c_void_p], _c_int),
    "kbner_crf_viterbi": ([_c_void_p] * 5 + [_c_int] * 6 + [_c_void_p] * 3, _c_int),
    "kbner_crf_nll_fwd": ([_c_void_p] * 5 + [_c_int] * 5 + [_c_void_p] * 5, _c_int),
    "kbner_crf_nll_bwd": ([_c_void_p] * 8 + [_c_int] * 5 + [_c_void_p] * 3, _c_int),
    "kbner_embed_ln_fwd": ([_c_void_p] * 6 + [_c_float, _c_int] + [_c_int] * 5 + [_c_void_p] * 2, _c_int),
    "kbner_layernorm_fwd": ([_c_void_p] * 3 + [_c_float, _c_int, _c_int] + [_c_void_p] * 4, _c_int),
    "kbner_gather_tagproj_fwd": ([_c_void_p] * 6 + [_c_int] * 5 + [_c_void_p] * 2, _c_int),
    "kbner_gemm_bf16_tn": ([_c_void_p] * 5 + [_c_int] * 7 + [_c_void_p], _c_int),
    "kbner_gemm_bf16": ([_c_void_p] * 6 + [_c_int] * 9 + [_c_void_p], _c_int),
    "kbner_gemm_wgrad_group": ([_c_int] + [_c_void_p] * 7 + [_c_int, _c_void_p], _c_int),
    "kbner_attention_fwd": ([_c_void_p] * 2 + [_c_int] * 3 + [_c_void_p] * 3, _c_int),
    "kbner_attention_bwd": ([_c_void_p] * 5 + [_c_int] * 3 + [_c_void_p] * 4, _c_int),
    "kbner_layernorm_bwd": ([_c_void_p] * 5 + [_c_int] * 2 + [_c_void_p] * 5, _c_int),
    "kbner_add_layernorm_fwd": ([_c_void_p] * 5 + [_c_float, _c_int, _c_int] + [_c_void_p] * 4 + [ctypes.c_uint32, _c_float, _c_void_p],
                                _c_int),
    "kbner_add_layernorm_bwd": ([_c_void_p] * 8 + [_c_int] * 2 + [_c_void_p] * 6 + [ctypes.c_uint32, _c_float, _c_void_p], _c_int),
    "kbner_dropout_apply": ([_c_void_p, _c_int, _c_int, _c_int, _c_void_p, ctypes.c_uint32, _c_float, _c_void_p], _c_int),
    "kbner_attention_fwd_dropout": ([_c_void_p] * 2 + [_c_int] * 3 + [_c_void_p] * 3 + [ctypes.c_uint32, _c_float, _c_void_p], _c_int),
    "kbner_attention_bwd_dropout": ([_c_void_p] * 5 + [_c_int] * 3 + [_c_void_p] * 4 + [ctypes.c_uint32, _c_float, _c_void_p],
                                    _c_int),
    "kbner_gemm_bias_resid_layernorm": ([_c_void_p] * 6 + [_c_float, _c_void_p] + [_c_int] * 5 + [_c_void_p], _c_int),
    "kbner_gemm_ln_resident_clusters": ([_c_int], _c_int),
    "kbner_gemm_ln_grid": ([_c_void_p] * 6 + [_c_float, _c_void_p] + [_c_int] * 5 + [_c_void_p, ctypes.c_size_t, _c_void_p], _c_int),
    "kbner_debug_gemm_ln_timeline": ([_c_void_p], _c_int),
    "kbner_gemm_ln_workspace_bytes": ([_c_int, _c_int], ctypes.c_size_t),
    "kbner_gemm_bias_resid_layernorm_ws": ([_c_void_p] * 6 + [_c_float, _c_void_p] + [_c_int] * 5 + [_c_void_p, ctypes.c_size_t, _c_void_p],
                                           _c_int),
    "kbner_embed_ln_fwd_ex": ([_c_void_p] * 6 + [_c_float, _c_int] + [_c_int] * 5 + [_c_void_p] * 2 + [_c_int, _c_void_p], _c_int),
    "kbner_add_layernorm_fwd_res32": ([_c_void_p] * 5 + [_c_float, _c_int, _c_int] + [_c_void_p] * 2 + [_c_int, _c_void_p], _c_int),
    "kbner_bias_gelu_split": ([_c_void_p] * 2 + [_c_int] * 2 + [_c_void_p] * 2, _c_int),
    "kbner_attention_fwd_ex": ([_c_void_p] * 2 + [_c_int] * 3 + [_c_void_p, _c_int] + [_c_void_p] * 4 +
                               [ctypes.c_uint32, _c_float, _c_void_p], _c_int),
    "kbner_attention_bwd_ex": ([_c_void_p] * 6 + [_c_int] * 3 + [_c_void_p] * 4 + [ctypes.c_uint32, _c_float, _c_void_p], _c_int),
    "kbner_gather_tagproj_fwd_f32": ([_c_void_p] * 6 + [_c_int] * 5 + [_c_void_p] * 2, _c_int),
    "kbner_colsum_bf16": ([_c_void_p, _c_int, _c_int, _c_void_p, _c_void_p], _c_int),
    "kbner_embed_ln_bwd": ([_c_void_p] * 5 + [_c_float] + [_c_int] * 4 + [_c_void_p] * 7, _c_int),
    "kbner_gather_tagproj_bwd": ([_c_void_p] * 6 + [_c_int] * 5 + [_c_void_p] * 4, _c_int),
    "kbner_sumsq_f32": ([_c_void_p, ctypes.c_size_t, _c_void_p, _c_void_p], _c_int),
    "kbner_clip_coef": ([_c_void_p, _c_float, _c_float, _c_void_p, _c_void_p], _c_int),
    "kbner_adamw_step": ([_c_void_p] * 4 + [ctypes.c_size_t] + [_c_float] * 5 + [_c_int, _c_void_p, _c_float, _c_void_p],
                         _c_int),
    "kbner_adamw_step_ex": ([_c_void_p] * 5 + [ctypes.c_size_t] + [_c_float] * 5 + [_c_int, _c_void_p, _c_float, _c_void_p,
                                                                          ctypes.c_size_t, _c_void_p], _c_int),
    "kbner_pack_bf16": ([_c_void_p, _c_void_p, ctypes.c_size_t, _c_float, _c_void_p], _c_int),
    "kbner_sumsq_bf16": ([_c_void_p, ctypes.c_size_t, _c_void_p, _c_void_p], _c_int),
    "kbner_mark_rows": ([_c_void_p, ctypes.c_size_t, _c_int, _c_void_p, _c_void_p], _c_int),
    "kbner_adamw_rows": ([_c_void_p] * 5 + [_c_int, _c_int] + [_c_float] * 5 + [_c_int, _c_void_p, _c_float, _c_void_p], _c_int),
    "kbner_sumsq_rows_det": ([_c_void_p, _c_void_p, _c_int, _c_int, _c_void_p, _c_int, _c_void_p, _c_void_p], _c_int),
    "kbner_zero_rows": ([_c_void_p, _c_void_p, _c_int, _c_int, _c_void_p], _c_int),
    "kbner_rows_gather_bf16": ([_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_void_p, _c_int, _c_void_p], _c_int),
    "kbner_rows_scatter_add_bf16": ([_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_void_p, _c_void_p], _c_int),
    "kbner_sumsq_det": ([_c_void_p, ctypes.c_size_t, _c_int, _c_void_p, _c_int, _c_void_p, _c_void_p], _c_int),
}

_lib = None


class KbnerError(RuntimeError):
    pass


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise KbnerError(
            "kbner_b200: %s not found -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU / PyTorch fallback for this path)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (argtypes, restype) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is missing
        fn.argtypes = argtypes
        fn.restype = restype
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().kbner_last_error()
        raise KbnerError("%s failed (code %d): %s" % (what, rc, msg.decode() if msg else "?"))


def launch_count():
    return int(load().kbner_launch_count())
