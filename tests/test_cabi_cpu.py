"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/synchformer_b200.h declares (no compute calls)."""
import ctypes
import os
import re

from synchformer_b200 import _lib, build

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(REPO, 'include', 'synchformer_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(sfb_[a-z0-9_]+)\s*\(', text)))


def test_library_builds_and_exports_every_declared_symbol():
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    declared = _declared_symbols()
    assert len(declared) >= 14
    for name in declared:
        assert hasattr(lib, name), f'{name} declared in the header but not exported'
    assert sorted(_lib.SIGNATURES) == declared, 'ctypes signatures and header disagree'


def test_host_only_entry_points():
    lib = _lib.load()
    assert lib.sfb_abi_version() == 1
    assert isinstance(lib.sfb_last_error(), bytes)


def test_argument_validation_happens_before_any_launch():
    lib = _lib.load()
    # null pointers / bad shapes are rejected on the host with SFB_E_INVALID (-1) and an error text
    rc = lib.sfb_gemm_bf16(None, 8, None, None, None, 0, None, 8, 1, 8, 8, 0, 0, None)
    assert rc == -1 and b'null' in lib.sfb_last_error()
    rc = lib.sfb_mel_frontend(None, None, 0, None)
    assert rc == -1
    d = _lib.AttnDesc()
    assert lib.sfb_attention(ctypes.byref(d), None) == -1


def test_sass_contains_blackwell_tensor_and_tma_instructions():
    """tcgen05.mma -> UTCHMMA, tcgen05.ld -> LDTM, TMA -> UTMALDG (B200_PROFILING.md 'What proves a Blackwell-native kernel')."""
    import shutil
    import subprocess
    cuobjdump = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'
    if not os.path.exists(cuobjdump):
        import pytest
        pytest.skip('cuobjdump not available')
    sass = subprocess.run([cuobjdump, '-sass', build.LIB], capture_output=True, text=True).stdout
    for mnemonic in ('UTCHMMA', 'LDTM', 'UTMALDG', 'HMMA'):
        assert mnemonic in sass, mnemonic


def test_training_entry_points_validate_then_reach_the_launch():
    """N3 entry points, called the way ops.py calls them but with fake (never dereferenced) device addresses on a box WITHOUT a GPU:
    a well-formed call gets through the host-side validation and fails only at the CUDA launch (SFB_E_CUDA = -2), a malformed one is
    rejected before (SFB_E_INVALID = -1 / SFB_E_UNSUPPORTED = -3).  Catches argument-order / ctypes-signature slips without hardware."""
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip('would launch kernels on fake addresses')
    lib = _lib.load()
    P = lambda k: ctypes.c_void_p(0x10000 * (k + 1))          # 64 KB-aligned fake addresses
    B, T, M = 2, 30, 60
    ok = [
        lib.sfb_dropout(P(0), P(1), P(2), 0, M * 768, 0.1, 12345678901234, 3, None),
        lib.sfb_dropout(P(0), None, P(2), 1, M * 768, 0.0, 0, 0, None),
        lib.sfb_gelu_fwd(P(0), P(1), M * 3072, None),
        lib.sfb_gelu_bwd(P(0), P(1), P(2), M * 3072, None),
        lib.sfb_transpose_bf16(P(0), 768, M, 768, P(1), 64, None),
        lib.sfb_colsum(P(0), 1, 768, M, 768, P(1), None, 0, None),
        lib.sfb_colsum(P(0), 0, 2304, 5000, 2304, P(1), P(2), 64 * 2304, None),
        lib.sfb_layernorm_bwd(P(0), 768, 16, 30, 1, P(1), 768, P(2), 1e-5, P(3), 768, 0, ctypes.c_void_p(0x90000), ctypes.c_void_p(0x90000 + 768 * 4),
                              P(5), lib.sfb_layernorm_bwd_workspace_floats(32), 32, None),
        lib.sfb_attention_train_fwd(P(0), P(1), P(2), B, T, 8, 96, 0.102, 0.1, 99, 1, None),
        lib.sfb_attention_train_bwd(P(0), P(1), P(2), P(3), P(4), P(5), B, T, 8, 96, 0.102, 0.1, 99, 1, None),
        lib.sfb_sync_head_bwd(P(0), T, P(1), P(2), 1e-5, P(3), P(4), B, 21, P(5), P(6), P(7), P(8), P(9), P(10), None),
    ]
    assert ok == [-2] * len(ok), (ok, lib.sfb_last_error())
    assert lib.sfb_layernorm_bwd_workspace_floats(32) == 4 * 2 * 768 and lib.sfb_layernorm_bwd_workspace_floats(10 ** 6) <= 2 * 148 * 2 * 768
    bad = [
        lib.sfb_dropout(P(0), None, P(2), 0, 6, 0.1, 0, 0, None),                                   # n % 4
        lib.sfb_dropout(P(0), None, P(2), 0, 8, 1.0, 0, 0, None),                                   # p outside [0, 1)
        lib.sfb_gelu_fwd(P(0), P(1), 12, None),                                                     # n % 8
        lib.sfb_transpose_bf16(P(0), 768, M, 768, P(1), 56, None),                                  # ld_out < R
        lib.sfb_colsum(P(0), 1, 768, M, 767, P(1), None, 0, None),                                  # odd N
        lib.sfb_layernorm_bwd(P(0), 768, 16, 30, 1, P(1), 768, P(2), 1e-5, P(3), 768, 0, P(4), P(6), P(5), 10 ** 6, 32, None),   # dbeta != dgamma + 768
        lib.sfb_layernorm_bwd(P(0), 768, 16, 30, 1, P(1), 768, P(2), 1e-5, P(3), 768, 0, ctypes.c_void_p(0x90000), ctypes.c_void_p(0x90000 + 3072),
                              P(5), 100, 32, None),                                                 # workspace too small
        lib.sfb_attention_train_fwd(P(0), P(1), P(2), B, T, 8, 80, 0.1, 0.1, 99, 1, None),          # head_dim
        lib.sfb_sync_head_bwd(P(0), T, P(1), P(2), 1e-5, P(3), P(4), B, 65, P(5), P(6), P(7), P(8), P(9), P(10), None),          # n_cls > 64
    ]
    assert bad == [-1] * len(bad), bad
    assert lib.sfb_attention_train_fwd(P(0), P(1), P(2), 1, 600, 8, 96, 0.1, 0.0, 0, 0, None) == -3   # T too long for shared memory
    assert b'shared memory' in lib.sfb_last_error()


def test_encoder_backward_entry_points_validate_then_reach_the_launch():
    """N1 entry points, same idea as above: well-formed calls fail only at the launch (-2) on a box without a GPU, malformed ones earlier."""
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip('would launch kernels on fake addresses')
    lib = _lib.load()
    A = 0x100000
    d = _lib.AttnDesc()
    D, TOK = 768, 1569
    row, seg = 3 * D, TOK * 3 * D
    d.q, d.k, d.v, d.out = A + row * 2, A + (row + D) * 2, A + (row + 2 * D) * 2, 2 * A
    d.k_prefix, d.v_prefix, d.prefix_outer = A + D * 2, A + 2 * D * 2, seg
    d.q_outer, d.q_inner, d.q_row = seg, 196 * row, row
    d.kv_outer, d.kv_inner, d.kv_row = seg, 196 * row, row
    d.o_outer, d.o_inner, d.o_row = TOK * D, 196 * D, D
    d.n_outer, d.n_inner, d.n_heads, d.head_dim, d.Lq, d.Lk, d.scale = 2, 8, 12, 64, 196, 196, 0.125
    P = lambda k: ctypes.c_void_p(A * (k + 3))
    assert lib.sfb_attention_bwd_stats_floats(ctypes.byref(d)) == 2 * 8 * 12 * 196 * 2
    assert lib.sfb_attention_bwd(ctypes.byref(d), P(0), P(1), P(2), P(3), P(4), P(5), None) == -2, lib.sfb_last_error()
    assert lib.sfb_attention_bwd(ctypes.byref(d), P(0), P(1), P(2), P(3), None, P(5), None) == -1       # prefix without dprefix
    d.Lk = 2000
    assert lib.sfb_attention_bwd(ctypes.byref(d), P(0), P(1), P(2), P(3), P(4), P(5), None) == -3       # does not fit in shared memory
    d.Lk, d.kv_row = 196, row + 1
    assert lib.sfb_attention_bwd(ctypes.byref(d), P(0), P(1), P(2), P(3), P(4), P(5), None) == -1       # unaligned rows
    assert lib.sfb_attention_bwd_global_query(P(0), seg, P(1), P(2), seg, row, P(3), P(4), TOK * D, P(5), P(6), P(7), P(8), P(9), 2, 12, 64, TOK, 0.125,
                                              None) == -2
    assert lib.sfb_attention_bwd_global_query(P(0), seg, P(1), P(2), seg, row, P(3), P(4), TOK * D, P(5), P(6), P(7), None, P(9), 2, 12, 96, TOK, 0.125,
                                              None) == -1
    assert lib.sfb_droppath(P(0), P(1), P(2), 0, 4 * 1569, 1569, 0.2, 7, 3, None) == -2
    assert lib.sfb_droppath(P(0), None, P(2), 1, 4 * 1569 + 1, 1569, 0.2, 7, 3, None) == -1             # rows not a multiple of the sample size
    assert lib.sfb_gather_rows_bf16(P(0), 768, P(1), 2 * 1568, 1568, 1569, 1, None) == -2
    assert lib.sfb_gather_rows_bf16(P(0), 700, P(1), 2 * 1568, 1568, 1569, 1, None) == -1
