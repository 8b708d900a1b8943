"""Bind `synchformer_b200.ops` to the CPU SIMT emulator build of the N3 kernels (tests/emu/_build/libsfb_emu.so).

TEST INFRASTRUCTURE ONLY.  `install(monkeypatch)` makes the REAL wrappers of ops.py (dropout, gelu_fwd / gelu_bwd, transpose_bf16,
colsum, layernorm_bwd, attention_train_fwd / _bwd, sync_head_bwd) call the REAL C-ABI entry points and the REAL kernel code, compiled
for the emulator, on CPU tensors.  The PTX-free kernels that were verified on hardware in round 1 (LayerNorm, token assembly, im2col, casts,
head) are emulated from their real sources too - which checks the emulator against kernels known to be right; only the tcgen05 / mma.sync
kernels (GEMM, attention forward) are served by the torch stand-ins of tests/fake_ops.py with real bf16 dtypes.
"""
import ctypes
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

EMULATED = ('sfb_dropout', 'sfb_gelu_fwd', 'sfb_gelu_bwd', 'sfb_transpose_bf16', 'sfb_colsum', 'sfb_layernorm_bwd_workspace_floats',
            'sfb_layernorm_bwd', 'sfb_attention_train_fwd', 'sfb_attention_train_bwd', 'sfb_sync_head_bwd', 'sfb_last_error',
            'sfb_attention_bwd_stats_floats', 'sfb_attention_bwd', 'sfb_attention_bwd_global_query', 'sfb_droppath', 'sfb_gather_rows_bf16',
            'sfb_cross_entropy', 'sfb_optim_chunk_elems', 'sfb_grad_sqnorm', 'sfb_adam_step',
            'sfb_mean_tokens', 'sfb_mean_tokens_bwd', 'sfb_l2_normalize', 'sfb_l2_normalize_bwd', 'sfb_contrastive_loss', 'sfb_contrastive_loss_bwd',
            # verified on the B200 in round 1 (no inline PTX): emulating them checks the emulator against kernels known to be right
            'sfb_layernorm', 'sfb_im2col_video', 'sfb_im2col_video_clip', 'sfb_video_tokens', 'sfb_im2col_ast', 'sfb_ast_tokens', 'sfb_sync_tokens',
            'sfb_sync_head', 'sfb_cast_f32_bf16', 'sfb_rowstats_cast',
            # the mma.sync / ldmatrix / cp.async attention kernels (their PTX is emulated instruction by instruction)
            'sfb_attention', 'sfb_attention_extra_supported', 'sfb_attention_merge_partials')
STAND_INS = ('require_cuda', 'gemm')          # the tcgen05 GEMM is not emulated: torch stand-in (hardware-verified in round 1)

_emu = None


def load():
    global _emu
    if _emu is None:
        from synchformer_b200 import _lib
        from emu import build_emu
        lib = ctypes.CDLL(build_emu.build())
        for name in EMULATED:
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = _lib.SIGNATURES[name]
        lib.emu_launch_count.restype = ctypes.c_long
        lib.emu_set_schedule.argtypes = [ctypes.c_int, ctypes.c_ulonglong]
        _emu = lib
    return _emu


def install(monkeypatch):
    import fake_ops
    from synchformer_b200 import _lib, ops
    lib = load()
    monkeypatch.setattr(_lib, '_lib', lib)                      # _lib.load() returns the cached handle
    monkeypatch.setattr(ops, '_stream', lambda t: None)
    fake_ops.install(monkeypatch, round_bf16=True, names=STAND_INS, real_dtypes=True)
    return lib
