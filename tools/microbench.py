"""Per-kernel timings at the benchmark shapes (B=64 x S=8 -> 512 segments), CUDA events, L2 flushed by working-set size.
Not a bench value: an optimisation aid.  Usage: python tools/microbench.py [n_seg]"""
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from synchformer_b200 import ops  # noqa: E402

D = 768


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    M = n * 1569
    dev = 'cuda'
    res = {}
    x = torch.randn(M, D, device=dev)
    ln = torch.randn(M, D, device=dev).bfloat16()
    qkv = torch.randn(M, 3 * D, device=dev).bfloat16()
    att = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
    hid = torch.empty(M, 4 * D, device=dev, dtype=torch.bfloat16)
    g, b = torch.ones(D, device=dev), torch.zeros(D, device=dev)

    def w(nn, kk):
        return (torch.randn(nn, kk, device=dev) * 0.02).bfloat16(), torch.zeros(nn, device=dev)

    for name, a, N, K, kw in [('qkv 768->2304', ln, 2304, 768, dict(out=qkv)), ('proj 768->768 +res f32', att, 768, 768, dict(out=x, residual=x, out_f32=True)),
                              ('fc1 768->3072 gelu', ln, 3072, 768, dict(out=hid, gelu=True)), ('fc2 3072->768 +res f32', hid, 768, 3072, dict(out=x, residual=x, out_f32=True)),
                              ('kv 768->1536', ln, 1536, 768, dict()), ('proj 768->768 bf16 out (no res)', att, 768, 768, dict())]:
        W, bias = w(N, K)
        ms = timeit(lambda: ops.gemm(a, W, bias, **kw))
        fl = 2.0 * M * N * K
        res['gemm ' + name] = {'ms': ms, 'TFLOPs': fl / ms / 1e9}
        if os.environ.get('SFB_MB_CUBLAS') == '1':
            # library yardstick for the same shape: plain cuBLAS bf16 GEMM through torch (bias only, no GELU / residual / fp32 output), so
            # it is a LOWER bound on what a library path would need for the fused op.  Measurement aid only, never on the product path.
            cub = timeit(lambda: torch.nn.functional.linear(a, W, bias.bfloat16()))
            res['gemm ' + name].update({'cublas_ms': cub, 'cublas_TFLOPs': fl / cub / 1e9})
    # LayerNorm-fused variants (EMIT_LN producer, LN_FOLD consumer) next to the unfused pair they replace
    xb = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
    st = torch.empty(M, D // 64, 2, device=dev)
    for name, a, N, K, kw in [('proj 768->768 +res f32 +emit_ln', att, 768, 768, dict(out=x, residual=x, out_f32=True, emit_ln=(xb, st))),
                              ('fc2 3072->768 +res f32 +emit_ln', hid, 768, 3072, dict(out=x, residual=x, out_f32=True, emit_ln=(xb, st)))]:
        W, bias = w(N, K)
        ms = timeit(lambda: ops.gemm(a, W, bias, **kw))
        res['gemm ' + name] = {'ms': ms, 'TFLOPs': 2.0 * M * N * K / ms / 1e9}
    for name, N, kw in [('qkv 768->2304 ln_fold', 2304, dict(out=qkv)), ('fc1 768->3072 gelu ln_fold', 3072, dict(out=hid, gelu=True))]:
        W, bias = w(N, 768)
        cs = W.float().sum(1).contiguous()
        ms = timeit(lambda: ops.gemm(xb, W, bias, ln_fold=(st, cs, 1e-6), **kw))
        res['gemm ' + name] = {'ms': ms, 'TFLOPs': 2.0 * M * N * 768 / ms / 1e9}
    if os.environ.get('SFB_MB_GEMM_ONLY') == '1':
        for k, v in res.items():
            print(f'{k:36s} ' + '  '.join(f'{kk}={vv:9.3f}' for kk, vv in v.items()))
        return
    ms = timeit(lambda: ops.layernorm(x, g, b, 1e-6, out=ln))
    res['layernorm'] = {'ms': ms, 'GBs': M * D * 6 / ms / 1e6}
    row, seg = 3 * D, 1569 * 3 * D
    qkv.copy_(torch.randn(M, 3 * D, device=dev))          # the GEMM timings above left arbitrary (possibly non-finite) values here
    ms = timeit(lambda: ops.attention(qkv, qkv[:, D:], qkv[:, 2 * D:], att, q_strides=(seg, 0, row), kv_strides=(seg, 0, row), o_strides=(1569 * D, 0, D),
                                      n_outer=n, n_inner=1, n_heads=12, head_dim=64, Lq=1, Lk=1569, scale=0.125))
    res['attn cls 1x1569'] = {'ms': ms, 'GBs': M * 2 * D * 2 / ms / 1e6}
    ms = timeit(lambda: ops.attention(qkv[1:], qkv[1:, D:], qkv[1:, 2 * D:], att[1:], q_strides=(seg, row, 196 * row), kv_strides=(seg, row, 196 * row),
                                      o_strides=(1569 * D, D, 196 * D), n_outer=n, n_inner=196, n_heads=12, head_dim=64, Lq=8, Lk=8, scale=0.125,
                                      k_prefix=qkv[:, D:], v_prefix=qkv[:, 2 * D:], prefix_outer=seg))
    res['attn time 8x9'] = {'ms': ms, 'GBs': M * 4 * D * 2 / ms / 1e6}
    ms = timeit(lambda: ops.attention(qkv[1:], qkv[1:, D:], qkv[1:, 2 * D:], att[1:], q_strides=(seg, 196 * row, row), kv_strides=(seg, 196 * row, row),
                                      o_strides=(1569 * D, 196 * D, D), n_outer=n, n_inner=8, n_heads=12, head_dim=64, Lq=196, Lk=196, scale=0.125,
                                      k_prefix=qkv[:, D:], v_prefix=qkv[:, 2 * D:], prefix_outer=seg))
    res['attn space 196x197'] = {'ms': ms, 'GBs': M * 4 * D * 2 / ms / 1e6, 'TFLOPs': n * 8 * 12 * 4.0 * 196 * 197 * 64 / ms / 1e9}
    vis = torch.randn(n, 16, 3, 224, 224, device=dev).half()
    ms = timeit(lambda: ops.im2col_video(vis))
    res['im2col video'] = {'ms': ms, 'GBs': vis.numel() * 4 / ms / 1e6}
    wave = torch.randn(n, 10240, device=dev)
    res['mel'] = {'ms': timeit(lambda: ops.mel_frontend(wave))}
    # small-M GEMMs (AST / sync / aggregator tails)
    for name, Ms, N, K in [('ast qkv', n * 74, 2304, 768), ('ast fc1', n * 74, 3072, 768), ('ast fc2', n * 74, 768, 3072), ('sync qkv', 64 * 114, 2304, 768),
                           ('agg l1', n * 8, 3072, 768)]:
        a = torch.randn(Ms, K, device=dev).bfloat16()
        W, bias = w(N, K)
        ms = timeit(lambda: ops.gemm(a, W, bias))
        res['gemm ' + name] = {'ms': ms, 'TFLOPs': 2.0 * Ms * N * K / ms / 1e9}
    qa = torch.randn(n * 74, 3 * D, device=dev).bfloat16()
    oa = torch.empty(n * 74, D, device=dev, dtype=torch.bfloat16)
    res['attn ast 74x74'] = {'ms': timeit(lambda: ops.attention(qa, qa[:, D:], qa[:, 2 * D:], oa, q_strides=(74 * row, 0, row), kv_strides=(74 * row, 0, row),
                                                               o_strides=(74 * D, 0, D), n_outer=n, n_inner=1, n_heads=12, head_dim=64, Lq=74, Lk=74, scale=0.125))}
    qs = torch.randn(64 * 114, 3 * D, device=dev).bfloat16()
    os_ = torch.empty(64 * 114, D, device=dev, dtype=torch.bfloat16)
    res['attn sync 114x114 hd96'] = {'ms': timeit(lambda: ops.attention(qs, qs[:, D:], qs[:, 2 * D:], os_, q_strides=(114 * row, 0, row), kv_strides=(114 * row, 0, row),
                                                                       o_strides=(114 * D, 0, D), n_outer=64, n_inner=1, n_heads=8, head_dim=96, Lq=114, Lk=114,
                                                                       scale=1 / math.sqrt(96)))}
    for k, v in res.items():
        print(f'{k:36s} ' + '  '.join(f'{kk}={vv:9.3f}' for kk, vv in v.items()))
    print(json.dumps(res))


if __name__ == '__main__':
    main()
