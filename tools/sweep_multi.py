"""BASELINE.json config 5 at N > 1: clips/s for a GLOBAL batch in {1, 8, 32, 64, 128, 256} x segments in {8, 14}, the flattened (clip, segment)
list sharded over the ranks (parallel.synchformer_forward_sharded: one all-gather of segment features, sync transformer on a clip range,
one tiny all-gather of logits) - strong scaling at fixed global work.  Launch:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/sweep_multi.py

Device-resident inputs, CUDA events, max over ranks.  A table for DESIGN.md / profiles, not the bench line."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from synchformer_b200 import model as M, ops, parallel, synth  # noqa: E402


def main():
    rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    out = []
    for S in (8, 14):
        model = M.build_synchformer(n_segments=S, state_dict=synth.synthetic_state_dict(1337, n_segments=S), device=dev)
        for B in (1, 8, 32, 64, 128, 256):
            s0, s1 = parallel.shard_range(B * S, world, rank)
            n = s1 - s0
            g = torch.Generator(device=dev).manual_seed(rank)
            vis = (torch.rand(max(n, 0), 16, 3, 224, 224, device=dev, generator=g) * 2 - 1).half()
            wave = torch.randn(max(n, 0), 10240, device=dev, generator=g) * 0.2

            def step():
                with torch.no_grad():
                    mel = ops.mel_frontend(wave).unsqueeze(1) if n > 0 else torch.empty((0, 1, 128, 66), device=dev)
                    return parallel.synchformer_forward_sharded(model, vis, mel, B, S)
            for _ in range(2):
                step()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            k = 3 if B >= 128 else 5
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(k):
                lg = step()
            b.record()
            torch.cuda.synchronize()
            ms = torch.tensor([a.elapsed_time(b) / k], device=dev)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            if rank == 0:
                out.append({'n_gpus': world, 'batch': B, 'segments': S, 'ms_per_step': float(ms), 'clips_per_s': B / float(ms) * 1e3,
                            'finite': bool(torch.isfinite(lg).all())})
                print(json.dumps(out[-1]), flush=True)
            del vis, wave
            torch.cuda.empty_cache()
    if rank == 0:
        print(json.dumps({'sweep': out}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
