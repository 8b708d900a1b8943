"""Experiment aid: the tcgen05 space-attention kernel under SFB_ATTN_PIPE (initial stagger of the two tile pipelines, 0 = lock-step order),
timing + correctness on a slice + phase stamps of CTA 0 (SFB_ATTN_DBG_PTR).  Not on the product path."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from synchformer_b200 import ops  # noqa: E402

D = 768
NAMES = ['mma:full', 'mma:S0', 'mma:x', 'mma:S1|x', 'mma:end', 't0:preS', 't0:S', 't0:max', 't0:P', 't0:O', 't0:free', 't0:st', '-', 't1:preS', 't1:S', 't1:max', 't1:P', 't1:O',
         't1:free', 't1:st', 't0:sts', 't0:arrive', 't1:sts', 't1:arrive']


def attn(qkv, att, n):
    row, seg = 3 * D, 1569 * 3 * D
    ops.attention(qkv[1:], qkv[1:, D:], qkv[1:, 2 * D:], att[1:], q_strides=(seg, 196 * row, row), kv_strides=(seg, 196 * row, row),
                  o_strides=(1569 * D, 196 * D, D), n_outer=n, n_inner=8, n_heads=12, head_dim=64, Lq=196, Lk=196, scale=0.125,
                  k_prefix=qkv[:, D:], v_prefix=qkv[:, 2 * D:], prefix_outer=seg)


def main():
    n = 512
    qkv = torch.randn(n * 1569, 3 * D, device='cuda').bfloat16()
    att = torch.empty(n * 1569, D, device='cuda', dtype=torch.bfloat16)
    t = qkv[:1569].float().view(1569, 3, 12, 64)
    q, k, v = t[:, 0].permute(1, 0, 2), t[:, 1].permute(1, 0, 2), t[:, 2].permute(1, 0, 2)
    kk, vv = torch.cat([k[:, :1], k[:, 1:197]], 1), torch.cat([v[:, :1], v[:, 1:197]], 1)
    ref = torch.softmax(q[:, 1:197] @ kk.transpose(-1, -2) * 0.125, -1) @ vv
    for pipe in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else '0,1500,2300,3000,0').split(',')]:
        os.environ['SFB_ATTN_PIPE'] = str(pipe)
        for _ in range(3):
            attn(qkv, att, n)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(10):
            attn(qkv, att, n)
        e.record()
        torch.cuda.synchronize()
        got = att[1:197].float().view(196, 12, 64).permute(1, 0, 2)
        last = att[-196:].float()
        print(f'SFB_ATTN_PIPE={pipe}: {s.elapsed_time(e) / 10:.3f} ms   rel-L2 vs torch {float((got - ref).norm() / ref.norm()):.2e}  finite {bool(torch.isfinite(last).all())}', flush=True)
    dbg = torch.zeros(12 * 24, device='cuda', dtype=torch.int64)
    os.environ['SFB_ATTN_DBG_PTR'] = str(dbg.data_ptr())
    for pipe in (3000,):
        os.environ['SFB_ATTN_PIPE'] = str(pipe)
        dbg.zero_()
        attn(qkv, att, 64)
        torch.cuda.synchronize()
        tt = dbg.cpu().view(12, 24)
        base = int(tt[4, 0])
        print(f'pipe {pipe}: clocks relative to mma:full of problem 4')
        for it in range(4, 9):
            print('  it', it, ' '.join(f'{nm}={int(tt[it, i]) - base}' for i, nm in enumerate(NAMES) if nm != '-'))
    os.environ.pop('SFB_ATTN_DBG_PTR')


if __name__ == '__main__':
    main()
