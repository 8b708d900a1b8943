"""Out-of-bounds check of the N3 kernels on the CPU SIMT emulator: every buffer handed to the C-ABI is placed flush against an
inaccessible guard page (its END for overruns, then its START for underruns), so a single element read or written outside a buffer
is a SIGSEGV here instead of an `illegal memory access` on the B200.  Run as a script (a crash must not take pytest down):

    python tests/emu/guarded_run.py        -> prints OK and exits 0

TEST INFRASTRUCTURE ONLY.
"""
import ctypes
import mmap
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from emu import binding  # noqa: E402

PAGE = mmap.PAGESIZE
libc = ctypes.CDLL(None, use_errno=True)
libc.mprotect.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
_keep = []


class Buf:
    """n_bytes of read/write memory with PROT_NONE pages on both sides; `front` puts the data right after the leading guard page,
    otherwise it ends exactly at the trailing guard page."""

    def __init__(self, n_bytes: int, front: bool, fill: int = 0x7F):
        assert n_bytes % 16 == 0 or not front
        pages = (n_bytes + PAGE - 1) // PAGE
        mm = mmap.mmap(-1, (pages + 2) * PAGE)
        _keep.append(mm)
        base = ctypes.addressof(ctypes.c_char.from_buffer(mm))
        ctypes.memset(base + PAGE, fill, pages * PAGE)
        for guard in (base, base + (pages + 1) * PAGE):
            if libc.mprotect(guard, PAGE, 0) != 0:
                raise OSError(ctypes.get_errno(), 'mprotect')
        self.addr = base + PAGE if front else base + (pages + 1) * PAGE - n_bytes
        self.n_bytes = n_bytes

    def ptr(self):
        return ctypes.c_void_p(self.addr)

    def np(self, dtype, shape):
        arr = np.ctypeslib.as_array((ctypes.c_uint8 * self.n_bytes).from_address(self.addr))
        return arr.view(dtype).reshape(shape)


def bf16_bytes(n):
    return ((n * 2 + 15) // 16) * 16


def run(front: bool):
    lib = binding.load()
    rng = np.random.default_rng(0)

    def f32(shape, scale=1.0):
        n = int(np.prod(shape))
        b = Buf(((n * 4 + 15) // 16) * 16, front)
        b.np(np.float32, -1)[:n] = rng.standard_normal(n).astype(np.float32) * scale
        return b

    def bf16(shape, scale=1.0):
        n = int(np.prod(shape))
        b = Buf(bf16_bytes(n), front)
        v = (rng.standard_normal(n).astype(np.float32) * scale).view(np.uint32) >> 16      # truncate to bf16
        b.np(np.uint16, -1)[:n] = v.astype(np.uint16)
        return b

    def out(n_bytes):
        return Buf(((n_bytes + 15) // 16) * 16, front)

    def ok(rc, what):
        if rc != 0:
            raise RuntimeError(f'{what}: rc {rc}: {lib.sfb_last_error().decode()}')

    # elementwise
    n = 333 * 768
    ok(lib.sfb_dropout(f32(n).ptr(), f32(n).ptr(), out(n * 4).ptr(), 0, n, 0.1, 77, 3, None), 'dropout f32')
    ok(lib.sfb_dropout(f32(n).ptr(), None, out(n * 2).ptr(), 1, n, 0.1, 77, 3, None), 'dropout bf16')
    n = 90 * 3072
    ok(lib.sfb_gelu_fwd(bf16(n).ptr(), out(n * 2).ptr(), n, None), 'gelu fwd')
    ok(lib.sfb_gelu_bwd(bf16(n).ptr(), bf16(n).ptr(), out(n * 2).ptr(), n, None), 'gelu bwd')
    # transpose: tails in both dimensions, padded leading dimension
    for R, C in ((90, 768), (594, 40), (7, 2304), (33, 33 * 8)):
        ld = (R + 7) // 8 * 8
        ok(lib.sfb_transpose_bf16(bf16(R * C).ptr(), C, R, C, out(C * ld * 2).ptr(), ld, None), f'transpose {R}x{C}')
    # column sums: N not a multiple of 64, M not a multiple of 8, both stages
    for M, N, is_bf in ((90, 768, 1), (1000, 2304, 1), (3, 45 * 768, 0), (513, 34, 0), (1, 768, 0)):
        src = bf16(M * N) if is_bf else f32(M * N)
        parts = min(64, (M + 63) // 64)
        ws = out(parts * N * 4) if parts > 1 else None
        ok(lib.sfb_colsum(src.ptr(), is_bf, N, M, N, out(N * 4).ptr(), ws.ptr() if ws else None, parts * N if ws else 0, None), f'colsum {M}x{N}')
    # LayerNorm backward: plain, and with the token gather of the sync sequence (B = 2, S = 3: T = 44)
    for rows, gather in ((90, None), (1000, None), (2 * 24, (24, 44, 1)), (2 * 18, (18, 44, 26))):
        group, stride, offset = gather if gather else (rows, rows, 0)
        n_dy = (rows // group - 1) * stride + offset + group
        nws = lib.sfb_layernorm_bwd_workspace_floats(rows)
        dgb = out(2 * 768 * 4)
        ok(lib.sfb_layernorm_bwd(f32(n_dy * 768).ptr(), 768, group, stride, offset, f32(rows * 768).ptr(), 768, f32(768).ptr(), 1e-5,
                                 f32(rows * 768).ptr(), 768, 1, ctypes.c_void_p(dgb.addr), ctypes.c_void_p(dgb.addr + 768 * 4), out(nws * 4).ptr(), nws,
                                 rows, None), f'layernorm bwd {rows}')
    # attention: T not a multiple of 32 / of 4, both head sizes
    for B, T, h, d in ((2, 45, 8, 96), (1, 198, 8, 96), (2, 74, 12, 64), (1, 33, 8, 96), (3, 1, 8, 96)):
        Dm = h * d
        qkv, o, lse = bf16(B * T * 3 * Dm, 0.5), out(B * T * Dm * 2), out(B * h * T * 4)
        ok(lib.sfb_attention_train_fwd(qkv.ptr(), o.ptr(), lse.ptr(), B, T, h, d, 0.1, 0.1, 5, 1, None), f'attention fwd T={T}')
        ok(lib.sfb_attention_train_bwd(qkv.ptr(), o.ptr(), bf16(B * T * Dm).ptr(), lse.ptr(), out(B * h * T * 4).ptr(), out(B * T * 3 * Dm * 2).ptr(), B, T, h, d,
                                       0.1, 0.1, 5, 1, None), f'attention bwd T={T}')
    # N1: strided attention backward with a shared prefix row (Motionformer layouts on one segment), global query, DropPath, gather
    from synchformer_b200._lib import AttnDesc
    D, TOK, h, d = 768, 1569, 12, 64
    row, seg = 3 * D, TOK * 3 * D
    # the mma.sync kernel (Lq >= 64, head_dim 64) on a 2-frame layout whose last key row ends at the guard page, then the CUDA-core pair (impl 1)
    for impl in (0, 1):
        rows_ = 2 * 196 + 1
        qkv, att, d_o = bf16(rows_ * 3 * D, 0.5), bf16(rows_ * D), bf16(rows_ * D)
        dqkv = out(rows_ * 3 * D * 2)
        desc = AttnDesc()
        desc.q, desc.k, desc.v, desc.out = qkv.addr + row * 2, qkv.addr + (row + D) * 2, qkv.addr + (row + 2 * D) * 2, att.addr + D * 2
        desc.k_prefix, desc.v_prefix, desc.prefix_outer = qkv.addr + D * 2, qkv.addr + 2 * D * 2, rows_ * row
        desc.q_outer, desc.q_inner, desc.q_row = rows_ * row, 196 * row, row
        desc.kv_outer, desc.kv_inner, desc.kv_row = rows_ * row, 196 * row, row
        desc.o_outer, desc.o_inner, desc.o_row = rows_ * D, 196 * D, D
        desc.n_outer, desc.n_inner, desc.n_heads, desc.head_dim, desc.Lq, desc.Lk, desc.scale, desc.impl = 1, 2, 12, 64, 196, 196, 0.125, impl
        desc.q_extra, desc.q_extra_outer, desc.extra_partial = None, 0, None
        ok(lib.sfb_attention_bwd(ctypes.byref(desc), ctypes.c_void_p(d_o.addr + D * 2), ctypes.c_void_p(dqkv.addr + row * 2),
                                 ctypes.c_void_p(dqkv.addr + (row + D) * 2), ctypes.c_void_p(dqkv.addr + (row + 2 * D) * 2), out(2 * 12 * 2 * 64 * 4).ptr(),
                                 out(lib.sfb_attention_bwd_stats_floats(ctypes.byref(desc)) * 4).ptr(), None), f'attention bwd space impl={impl}')
    for n, mode in ((1, 'time'),):
        qkv, att, d_o = bf16(n * TOK * 3 * D, 0.5), bf16(n * TOK * D), bf16(n * TOK * D)
        dqkv = out(n * TOK * 3 * D * 2)
        ctypes.memset(dqkv.addr, 0, dqkv.n_bytes)
        desc = AttnDesc()
        e = 2                                                       # bytes per element
        desc.q, desc.k, desc.v, desc.out = qkv.addr + row * e, qkv.addr + (row + D) * e, qkv.addr + (row + 2 * D) * e, att.addr + D * e
        desc.k_prefix, desc.v_prefix, desc.prefix_outer = qkv.addr + D * e, qkv.addr + 2 * D * e, seg
        if mode == 'time':
            st, so, desc.n_inner, desc.Lq, desc.Lk = (seg, row, 196 * row), (TOK * D, D, 196 * D), 196, 8, 8
        else:
            st, so, desc.n_inner, desc.Lq, desc.Lk = (seg, 196 * row, row), (TOK * D, 196 * D, D), 8, 196, 196
        desc.q_outer, desc.q_inner, desc.q_row = st
        desc.kv_outer, desc.kv_inner, desc.kv_row = st
        desc.o_outer, desc.o_inner, desc.o_row = so
        desc.n_outer, desc.n_heads, desc.head_dim, desc.scale, desc.impl = n, h, d, 0.125, 0
        desc.q_extra, desc.q_extra_outer, desc.extra_partial = None, 0, None
        n_stats = lib.sfb_attention_bwd_stats_floats(ctypes.byref(desc))
        part = out(desc.n_inner * n * h * 2 * d * 4)
        ok(lib.sfb_attention_bwd(ctypes.byref(desc), ctypes.c_void_p(d_o.addr + D * e), ctypes.c_void_p(dqkv.addr + row * e),
                                 ctypes.c_void_p(dqkv.addr + (row + D) * e), ctypes.c_void_p(dqkv.addr + (row + 2 * D) * e), part.ptr(),
                                 out(n_stats * 4).ptr(), None), f'attention bwd {mode}')
        pg = out(n * h * 2 * d * 4)
        ok(lib.sfb_colsum(part.ptr(), 0, n * h * 2 * d, desc.n_inner, n * h * 2 * d, pg.ptr(), None, 0, None), 'prefix colsum')
        ok(lib.sfb_attention_bwd_global_query(qkv.ptr(), seg, ctypes.c_void_p(qkv.addr + D * e), ctypes.c_void_p(qkv.addr + 2 * D * e), seg, row, att.ptr(),
                                              d_o.ptr(), TOK * D, dqkv.ptr(), ctypes.c_void_p(dqkv.addr + D * e), ctypes.c_void_p(dqkv.addr + 2 * D * e),
                                              pg.ptr(), out(n * h * TOK * 2 * 4).ptr(), n, h, d, TOK, 0.125, None), f'global query {mode}')
    # aggregator-style: one query per group, strided keys (AST frequency aggregator: 6 groups of 12 keys in 72 rows), shared prefix
    n = 3
    kv, ao, dao, qrep = bf16(n * 72 * 1536, 0.5), bf16(n * 6 * D), bf16(n * 6 * D), bf16(n * 6 * D, 0.5)
    cls_qkv = bf16(3 * D, 0.5)
    desc = AttnDesc()
    desc.q, desc.k, desc.v, desc.out = qrep.addr, kv.addr, kv.addr + D * 2, ao.addr
    desc.k_prefix, desc.v_prefix, desc.prefix_outer = cls_qkv.addr + D * 2, cls_qkv.addr + 2 * D * 2, 0
    desc.q_outer, desc.q_inner, desc.q_row = 6 * D, D, D
    desc.kv_outer, desc.kv_inner, desc.kv_row = 72 * 1536, 1536, 6 * 1536
    desc.o_outer, desc.o_inner, desc.o_row = 6 * D, D, D
    desc.n_outer, desc.n_inner, desc.n_heads, desc.head_dim, desc.Lq, desc.Lk, desc.scale, desc.impl = n, 6, 12, 64, 1, 12, 0.125, 0
    desc.q_extra, desc.q_extra_outer, desc.extra_partial = None, 0, None
    dkv = out(n * 72 * 1536 * 2)
    ok(lib.sfb_attention_bwd(ctypes.byref(desc), dao.ptr(), out(n * 6 * D * 2).ptr(), dkv.ptr(), ctypes.c_void_p(dkv.addr + D * 2), out(6 * n * 12 * 2 * 64 * 4).ptr(),
                             out(lib.sfb_attention_bwd_stats_floats(ctypes.byref(desc)) * 4).ptr(), None), 'attention bwd aggregator')
    rows = 4 * 1569
    ok(lib.sfb_droppath(f32(rows * 768).ptr(), f32(rows * 768).ptr(), out(rows * 768 * 4).ptr(), 0, rows, 1569, 0.3, 5, 2, None), 'droppath')
    ok(lib.sfb_droppath(f32(rows * 768).ptr(), None, out(rows * 768 * 2).ptr(), 1, rows, 1569, 0.3, 5, 2, None), 'droppath bf16')
    ok(lib.sfb_gather_rows_bf16(f32(rows * 768).ptr(), 768, out(4 * 1568 * 768 * 2).ptr(), 4 * 1568, 1568, 1569, 1, None), 'gather rows')
    # head
    for B, T, n_cls in ((2, 44, 21), (5, 198, 2), (1, 30, 64)):
        ok(lib.sfb_sync_head_bwd(f32(B * T * 768).ptr(), T, f32(768).ptr(), f32(768).ptr(), 1e-5, f32(n_cls * 768, 0.03).ptr(), f32(B * n_cls).ptr(), B, n_cls,
                                 out(B * T * 768 * 4).ptr(), out(768 * 4).ptr(), out(768 * 4).ptr(), out(n_cls * 768 * 4).ptr(), out(((n_cls * 4 + 15) // 16) * 16).ptr(),
                                 out(B * 3 * 768 * 4).ptr(), None), f'head bwd B={B}')


def selftest_must_crash():
    """a deliberately short output buffer: the guard page has to turn the overrun into a SIGSEGV"""
    lib = binding.load()
    R, C = 64, 64
    src = Buf(R * C * 2, False)
    dst = Buf((C - 1) * R * 2, False)                       # one row short
    lib.sfb_transpose_bf16(src.ptr(), C, R, C, dst.ptr(), R, None)


if __name__ == '__main__':
    if '--selftest' in sys.argv:
        selftest_must_crash()
        print('NOT CAUGHT')
        sys.exit(0)
    run(front=False)      # overruns
    run(front=True)       # underruns
    print('OK')
