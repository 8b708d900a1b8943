"""Context number, NOT a bench value: the torch-eager restatement of the reference forward (the oracle's op sequence, i.e. the
same cuBLAS / ATen library kernels the reference dispatches to) timed on the GPU under bf16 / fp16 autocast and fp32.
North-star's ">= 10x the reference single-GPU PyTorch clips/sec" needs a denominator; the Python reference itself cannot
travel to the GPU box, the restatement can.  Test infrastructure (imports oracle/)."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synchformer_oracle as O  # noqa: E402
from synchformer_b200 import synth  # noqa: E402


def main():
    S = 8
    dev = 'cuda'
    sd = {k: v.to(dev) for k, v in synth.synthetic_state_dict(1337, n_segments=S).items()}
    torch.backends.cudnn.benchmark = True
    res = {}
    for name, B, ctx in [('bf16_autocast', 16, torch.autocast('cuda', dtype=torch.bfloat16)),
                         ('fp16_autocast', 16, torch.autocast('cuda', dtype=torch.float16)),
                         ('fp32_tf32', 8, None)]:
        torch.backends.cuda.matmul.allow_tf32 = True
        vis = (torch.rand(B, S, 16, 3, 224, 224, device=dev) * 2 - 1)
        aud = torch.randn(B, S, 1, 128, 66, device=dev)
        def run():
            with torch.no_grad():
                if ctx is None:
                    return O.forward(sd, vis, aud, device=dev)[1]
                with ctx:
                    return O.forward(sd, vis, aud, device=dev)[1]
        try:
            for _ in range(2):
                run()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            n = 3
            for _ in range(n):
                run()
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / n
            res[name] = {'batch': B, 'segments': S, 'ms_per_step': dt * 1e3, 'clips_per_s': B / dt,
                         'peak_mem_gb': torch.cuda.max_memory_allocated() / 2 ** 30}
        except Exception as e:  # OOM etc.
            res[name] = {'error': repr(e)[:200]}
        del vis, aud
        torch.cuda.empty_cache()
    print(json.dumps(res, indent=1))


if __name__ == '__main__':
    main()
