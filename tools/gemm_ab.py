"""Same-process, interleaved A/B of the fp32-output GEMM epilogues at the benchmark shapes (512 segments): per-thread epilogue
(SFB_GEMM_F32_TMA=0) vs the TMA epilogue.  The switch is read per call,
so the variants alternate round-robin inside one process (same clocks, same thermal state); medians over the rounds are printed."""
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from synchformer_b200 import ops  # noqa: E402

D = 768
VARIANTS = {'per-thread': {'SFB_GEMM_F32_TMA': '0'}, 'tma': {'SFB_GEMM_F32_TMA': '1'}}


def timeit(fn, iters=40):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 7
    M = n * 1569
    dev = 'cuda'
    x = torch.randn(M, D, device=dev)
    att = torch.randn(M, D, device=dev).bfloat16()
    hid = torch.randn(M, 4 * D, device=dev).bfloat16()
    xb = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
    st = torch.empty(M, D // 64, 2, device=dev)
    wp = (torch.randn(D, D, device=dev) * 0.02).bfloat16()
    w2 = (torch.randn(D, 4 * D, device=dev) * 0.01).bfloat16()
    bias = torch.zeros(D, device=dev)
    cases = {'proj +res': lambda: ops.gemm(att, wp, bias, out=x, residual=x, out_f32=True),
             'fc2 +res': lambda: ops.gemm(hid, w2, bias, out=x, residual=x, out_f32=True),
             'proj +res +emit_ln': lambda: ops.gemm(att, wp, bias, out=x, residual=x, out_f32=True, emit_ln=(xb, st)),
             'fc2 +res +emit_ln': lambda: ops.gemm(hid, w2, bias, out=x, residual=x, out_f32=True, emit_ln=(xb, st)),
             'patch-embed style (fp32 out, no residual)': lambda: ops.gemm(att, wp, bias, out=x, out_f32=True)}
    res = {c: {v: [] for v in VARIANTS} for c in cases}
    for _ in range(rounds):
        for c, fn in cases.items():
            for v, env in VARIANTS.items():
                os.environ.update(env)
                res[c][v].append(timeit(fn))
            x.normal_()                                  # the in-place residual adds must not run away
    for c in cases:
        print(f'{c:44s}' + '  '.join(f'{v} {statistics.median(res[c][v]):.3f} ms' for v in VARIANTS))
    # LN_FOLD consumers next to their plain versions (same interleaving; the statistics / bf16 copy come from the EMIT_LN run above)
    os.environ['SFB_GEMM_F32_TMA'] = '1'
    ln = torch.randn(M, D, device=dev).bfloat16()
    qkv = torch.empty(M, 3 * D, device=dev, dtype=torch.bfloat16)
    ops.gemm(att, wp, bias, out=x, residual=x, out_f32=True, emit_ln=(xb, st))
    wq = (torch.randn(3 * D, D, device=dev) * 0.02).bfloat16()
    w1 = (torch.randn(4 * D, D, device=dev) * 0.02).bfloat16()
    bq, b1 = torch.zeros(3 * D, device=dev), torch.zeros(4 * D, device=dev)
    csq, cs1 = wq.float().sum(1).contiguous(), w1.float().sum(1).contiguous()
    pairs = {'qkv': (lambda: ops.gemm(ln, wq, bq, out=qkv), lambda: ops.gemm(xb, wq, bq, out=qkv, ln_fold=(st, csq, 1e-6))),
             'fc1 + gelu': (lambda: ops.gemm(ln, w1, b1, out=hid, gelu=True), lambda: ops.gemm(xb, w1, b1, out=hid, gelu=True, ln_fold=(st, cs1, 1e-6)))}
    for name, (plain, fold) in pairs.items():
        tp, tf = [], []
        for _ in range(rounds):
            tp.append(timeit(plain)), tf.append(timeit(fold))
        print(f'{name:44s}plain {statistics.median(tp):.3f} ms  ln_fold {statistics.median(tf):.3f} ms')


if __name__ == '__main__':
    main()
