"""ctypes binding of `libsynchformer_b200.so` (the C-ABI declared in include/synchformer_b200.h).

There is no CPU or eager-PyTorch fallback: if the library is missing or fails to load, importing the ops raises.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_uint32, c_uint64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'lib', 'libsynchformer_b200.so')

SFB_GEMM_GELU, SFB_GEMM_RESIDUAL, SFB_GEMM_OUT_F32, SFB_GEMM_EMIT_LN, SFB_GEMM_LN_FOLD = 1, 2, 4, 8, 16


class AttnDesc(Structure):
    _fields_ = [
        ('q', c_void_p), ('k', c_void_p), ('v', c_void_p),
        ('k_prefix', c_void_p), ('v_prefix', c_void_p),
        ('out', c_void_p),
        ('q_outer', c_int64), ('q_inner', c_int64), ('q_row', c_int64),
        ('kv_outer', c_int64), ('kv_inner', c_int64), ('kv_row', c_int64),
        ('o_outer', c_int64), ('o_inner', c_int64), ('o_row', c_int64),
        ('prefix_outer', c_int64),
        ('n_outer', c_int32), ('n_inner', c_int32), ('n_heads', c_int32), ('head_dim', c_int32),
        ('Lq', c_int32), ('Lk', c_int32),
        ('scale', c_float),
        ('impl', c_int32),
        ('q_extra', c_void_p), ('q_extra_outer', c_int64), ('extra_partial', c_void_p),
    ]


# name -> (restype, argtypes); must list EVERY symbol declared in include/synchformer_b200.h
SIGNATURES = {
    'sfb_abi_version': (c_int, []),
    'sfb_last_error': (c_char_p, []),
    'sfb_device_check': (c_int, []),
    'sfb_gemm_bf16': (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64,
                              c_int, c_int, c_int, c_int, c_int, c_void_p]),
    'sfb_gemm_bf16_ln': (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64,
                                 c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_float, c_void_p, c_void_p, c_int64, c_void_p]),
    'sfb_rowstats_cast': (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int, c_void_p]),
    'sfb_layernorm': (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_float,
                              c_void_p, c_void_p, c_float, c_int, c_int, c_int, c_int, c_void_p]),
    'sfb_attention': (c_int, [POINTER(AttnDesc), c_void_p]),
    'sfb_attention_extra_supported': (c_int, [POINTER(AttnDesc)]),
    'sfb_attention_merge_partials': (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_int, c_void_p]),
    'sfb_im2col_video': (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p]),
    'sfb_im2col_video_clip': (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    'sfb_video_tokens': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    'sfb_im2col_ast': (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    'sfb_ast_tokens': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    'sfb_sync_tokens': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    'sfb_sync_head': (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p,
                              c_int, c_int, c_void_p]),
    'sfb_cast_f32_bf16': (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    'sfb_mel_frontend': (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    'sfb_mel_frontend_clip': (c_int, [c_void_p, c_int64, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    # N3: training step of the synchronisation module
    'sfb_dropout': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_float, c_uint64, c_uint32, c_void_p]),
    'sfb_gelu_fwd': (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    'sfb_gelu_bwd': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    'sfb_transpose_bf16': (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_int64, c_void_p]),
    'sfb_colsum': (c_int, [c_void_p, c_int, c_int64, c_int, c_int, c_void_p, c_void_p, c_int64, c_void_p]),
    'sfb_layernorm_bwd_workspace_floats': (c_int, [c_int]),
    'sfb_layernorm_bwd': (c_int, [c_void_p, c_int64, c_int, c_int, c_int, c_void_p, c_int64, c_void_p, c_float, c_void_p, c_int64, c_int,
                                  c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    'sfb_attention_train_fwd': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_float, c_uint64, c_uint32,
                                        c_void_p]),
    'sfb_attention_train_bwd': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float,
                                        c_float, c_uint64, c_uint32, c_void_p]),
    'sfb_sync_head_bwd': (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    'sfb_cross_entropy': (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    'sfb_optim_chunk_elems': (c_int, []),
    'sfb_grad_sqnorm': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    'sfb_adam_step': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_float, c_float, c_float, c_float, c_float, c_float,
                              c_float, c_void_p]),
    # N1: backward of the encoders
    'sfb_attention_bwd_stats_floats': (c_int64, [POINTER(AttnDesc)]),
    'sfb_attention_bwd': (c_int, [POINTER(AttnDesc), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    'sfb_attention_bwd_global_query': (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_int64, c_void_p,
                                               c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p]),
    'sfb_droppath': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int, c_float, c_uint64, c_uint32, c_void_p]),
    'sfb_gather_rows_bf16': (c_int, [c_void_p, c_int64, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    # N1: contrastive tail
    'sfb_mean_tokens': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    'sfb_mean_tokens_bwd': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    'sfb_l2_normalize': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    'sfb_l2_normalize_bwd': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    'sfb_contrastive_loss': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_void_p]),
    'sfb_contrastive_loss_bwd': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                         c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
}

_lib = None


def load() -> ctypes.CDLL:
    """Load the shared library (building is `python -m synchformer_b200.build` / `__graft_entry__.build()`)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f'{LIB_PATH} not found: the CUDA extension is required (no CPU fallback). '
                f'Build it with `python -m synchformer_b200.build`.')
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError if a declared symbol is missing
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


class SfbError(RuntimeError):
    pass


def check(rc: int, what: str):
    if rc != 0:
        msg = load().sfb_last_error()
        raise SfbError(f'{what} failed with code {rc}: {msg.decode() if msg else ""}')
