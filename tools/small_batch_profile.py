"""One clip (B = 1, S = 14 by default) through the whole forward, for a per-kernel duration list:

    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/small_launches.csv \
        python tools/small_batch_profile.py 1 14

Only ONE forward (after warm-up) lies inside the cudaProfilerStart/Stop range.  Prints the CUDA-event time of the same forward without the
profiler when run plainly.  Measurement aid: where do the ~9 ms of a single clip go - kernel time or the gaps between ~290 launches?"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from synchformer_b200 import model as M, ops, synth  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    S = int(sys.argv[2]) if len(sys.argv) > 2 else 14
    dev = torch.device('cuda', 0)
    model = M.build_synchformer(n_segments=S, state_dict=synth.synthetic_state_dict(1337, n_segments=S), device=dev)
    g = torch.Generator(device=dev).manual_seed(0)
    vis = (torch.rand(B, S, 16, 3, 224, 224, device=dev, generator=g) * 2 - 1).half()
    wave = torch.randn(B * S, 10240, device=dev, generator=g) * 0.2

    def step():
        with torch.no_grad():
            mel = ops.mel_frontend(wave).view(B, S, 1, 128, 66)
            return model(vis, mel)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        step()
    b.record()
    torch.cuda.synchronize()
    print(f'B={B} S={S}: {a.elapsed_time(b) / 10:.3f} ms per forward (CUDA events, 10 forwards)')
    torch.cuda.cudart().cudaProfilerStart()
    step()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()


if __name__ == '__main__':
    main()
