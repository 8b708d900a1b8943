"""Library yardstick under ncu (measurement aid, never on the product path): torch.nn.functional.linear (cuBLAS / cuBLASLt bf16) on one of the
benchmark GEMM shapes, so that its kernel name (tile / cluster shape), tensor-pipe %, L2 and shared-memory traffic can be read next to
ours.      ncu --set full -k regex:'nvjet|cutlass|gemm|xmma' -s 2 -c 1 -o out python tools/cublas_probe.py qkv"""
import sys

import torch

SHAPES = {'qkv': (512 * 1569, 2304, 768), 'fc1': (512 * 1569, 3072, 768), 'fc2': (512 * 1569, 768, 3072), 'proj': (512 * 1569, 768, 768)}


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else 'qkv'
    M, N, K = SHAPES[name]
    a = (torch.randn(M, K, device='cuda') * 1.0).bfloat16()
    w = (torch.randn(N, K, device='cuda') * 0.02).bfloat16()
    for _ in range(4):
        out = torch.nn.functional.linear(a, w)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        out = torch.nn.functional.linear(a, w)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 5
    print(f'cublas {name} M={M} N={N} K={K}: {ms:.3f} ms = {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s (plain bf16 GEMM, no bias / GELU / fp32 out)', float(out[0, 0]))


if __name__ == '__main__':
    main()
