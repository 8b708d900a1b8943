"""BASELINE.json config 5 at N = 1: clips/s over batch in {1, 8, 32, 128, 256} x segments in {8, 14} (device-resident inputs, CUDA events).
Not the bench line (that is config 2); a table for DESIGN.md / profiles."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from synchformer_b200 import model as M, ops, synth  # noqa: E402


def main():
    dev = torch.device('cuda:0')
    out = []
    for S in (8, 14):
        model = M.build_synchformer(n_segments=S, state_dict=synth.synthetic_state_dict(1337, n_segments=S), device=dev)
        for B in (1, 8, 32, 128, 256):
            g = torch.Generator(device=dev).manual_seed(0)
            vis = (torch.rand(B, S, 16, 3, 224, 224, device=dev, generator=g) * 2 - 1).half()
            wave = torch.randn(B, S, 10240, device=dev, generator=g) * 0.2

            def step():
                with torch.no_grad():
                    return model(vis, ops.mel_frontend(wave).unsqueeze(2))[1]
            for _ in range(2):
                step()
            torch.cuda.synchronize()
            n = 3 if B >= 128 else 5
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(n):
                lg = step()
            e.record()
            torch.cuda.synchronize()
            ms = s.elapsed_time(e) / n
            out.append({'batch': B, 'segments': S, 'ms_per_step': ms, 'clips_per_s': B / ms * 1e3, 'finite': bool(torch.isfinite(lg).all())})
            print(json.dumps(out[-1]), flush=True)
            del vis, wave
            torch.cuda.empty_cache()
    print(json.dumps({'sweep': out}))


if __name__ == '__main__':
    main()
