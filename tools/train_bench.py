#!/usr/bin/env python
"""Training-step benchmark (BASELINE.json configs[3]: sync.yaml training step, fwd + bwd + DDP all-reduce, bf16).

    python tools/train_bench.py [--batch 32] [--segments 14] [--steps 5] [--warmup 3]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/train_bench.py --batch 32

One step = what scripts/train_sync.py:177-183 + scripts/train_utils.py:373-386 do per iteration with the extractors frozen
(configs/sync.yaml:8,20): forward of the frozen encoders (no grad), vproj / aproj + sync transformer forward with dropout 0.1, cross
entropy, GradScaler-scaled backward through the hand-written kernels, DDP all-reduce of the 22.6 M gradients (N > 1), unscale,
clip_grad_norm_(1), Adam(lr 2e-6 x N, eps 1e-7) step.  Inputs are device-resident synthetic tensors; weak scaling (`--batch` clips per
rank).  Prints one JSON line on rank 0: clips/s (all ranks), ms/step (max over ranks, CUDA events), the split encoder-forward /
sync-module forward+backward / optimizer, kernel launches of this library per step.

A measurement helper for SURVEY.md §8f N3, not the driver's bench (bench.py keeps the inference metric).
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=32, help='clips per GPU (config 4: 256 over 8 GPUs)')
    ap.add_argument('--segments', type=int, default=14)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--mode', default='sync', choices=['sync', 'avclip', 'avclip_fwd'], help="sync: BASELINE config 4 (stage II, frozen extractors); "
                    "avclip: stage-I contrastive step (both encoders train, SURVEY.md 8f N1; --batch x --segments segments per GPU)")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    from synchformer_b200 import model as M, ops, synth
    ops.device_check()
    B, S = args.batch, args.segments
    torch.manual_seed(1337 + rank)
    if args.mode == 'avclip':
        return avclip_bench(args, rank, world, local, dev)
    if args.mode == 'avclip_fwd':
        return avclip_forward_bench(args, rank, world, local, dev)
    model = M.build_synchformer(n_segments=S, state_dict=synth.synthetic_state_dict(1337, n_segments=S), device=dev)
    for ext in (model.vfeat_extractor, model.afeat_extractor):            # get_model: is_trainable False (train_utils.py:199-204)
        ext.requires_grad_(False)

    class SyncHalf(torch.nn.Module):
        """The differentiable half of Synchformer.forward (sync_model.py:55-70); the frozen encoders run outside, under no_grad, so that
        their time is reported separately.  DDP (train_utils.py:205-210) reduces exactly the gradients this module produces."""

        def __init__(self, m):
            super().__init__()
            self.m = m

        def forward(self, vf, af, tgt):
            v, a = self.m.project(vf, af)
            logits = self.m.transformer(v, a)
            return self.m.compute_loss(logits, tgt), logits

    half = SyncHalf(model)
    net = torch.nn.parallel.DistributedDataParallel(half, device_ids=[local]) if world > 1 else half
    opt = torch.optim.Adam(model.parameters(), 2e-6 * world, (0.9, 0.999), 1e-7, 0)        # get_optimizer train_utils.py:217-228
    scaler = torch.amp.GradScaler('cuda')

    g = torch.Generator(device=dev).manual_seed(rank)
    vis = (torch.rand((B, S, 16, 3, 224, 224), device=dev, generator=g, dtype=torch.float16) - 0.5) / 0.5
    aud = torch.randn((B, S, 1, 128, 66), device=dev, generator=g)
    targets = torch.randint(0, 21, (B,), device=dev, generator=g)

    def toggle_train():                                                   # toggle_mode train_utils.py:330-342
        model.train()
        model.vfeat_extractor.eval()
        model.afeat_extractor.eval()

    ev = {k: [] for k in ('fwd_enc', 'sync_fwd_bwd', 'optim')}

    def step(record: bool):
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        opt.zero_grad(set_to_none=True)
        marks[0].record()
        with torch.no_grad():                                             # what Synchformer.forward does for the frozen extractors
            vf = model.extract_vfeats(vis)
            af = model.extract_afeats(aud)
        marks[1].record()
        loss, _ = net(vf, af, targets)
        scaler.scale(loss).backward()
        marks[2].record()
        scaler.unscale_(opt)
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        scaler.step(opt)
        scaler.update()
        marks[3].record()
        if record:
            torch.cuda.synchronize()
            for k, (i, j) in zip(ev, ((0, 1), (1, 2), (2, 3))):
                ev[k].append(marks[i].elapsed_time(marks[j]))
        return loss

    toggle_train()
    for _ in range(args.warmup):
        step(False)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    n0 = ops.launch_count()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        loss = step(True)
    t1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([t0.elapsed_time(t1) / args.steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.barrier()
    if rank == 0:
        mean = lambda xs: sum(xs) / max(1, len(xs))
        print(json.dumps({
            'metric': 'clips/sec training step (sync.yaml, frozen extractors, fwd + bwd + optimizer)', 'value': B * world / (float(ms) / 1e3),
            'unit': 'clips/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': float(ms), 'dtype': 'bf16',
            'data': 'synthetic', 'scaling': 'weak', 'config': {'workload': f'sync.yaml training step, batch={B} clips/GPU, {S} segments/clip',
                                                               'optimizer': 'torch.optim.Adam + GradScaler + clip_grad_norm_(1)'},
            'split_ms': {k: mean(v) for k, v in ev.items()}, 'gpu_launches_per_step': (ops.launch_count() - n0) / args.steps,
            'loss': float(loss)
        }), flush=True)
    if world > 1:
        dist.destroy_process_group()


def avclip_bench(args, rank, world, local, dev):
    """Stage I: AVCLIP forward (DropPath) + backward of both encoders + AdamW (training/train.py:122-154), per-rank negatives."""
    from synchformer_b200 import avclip, ops, synth
    B, S = args.batch, args.segments
    model = avclip.AVCLIP().to(dev)
    sd = synth.synthetic_state_dict(1337, n_segments=S)
    model.v_encoder.load_state_dict({k[len('vfeat_extractor.'):]: v for k, v in sd.items() if k.startswith('vfeat_extractor.')})
    model.a_encoder.load_state_dict({k[len('afeat_extractor.'):]: v for k, v in sd.items() if k.startswith('afeat_extractor.')})
    model.train()
    net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local]) if world > 1 else model
    opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0)
    g = torch.Generator(device=dev).manual_seed(rank)
    rgb = (torch.rand((B, S, 3, 16, 224, 224), device=dev, generator=g, dtype=torch.float16) - 0.5) / 0.5
    aud = torch.randn((B, S, 66, 128), device=dev, generator=g)
    ev = {'fwd': [], 'bwd': [], 'optim': []}

    def step(record):
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        opt.zero_grad(set_to_none=True)
        marks[0].record()
        loss = net(rgb, aud)['losses']['segment_contrastive_loss']
        marks[1].record()
        loss.backward()
        marks[2].record()
        opt.step()
        marks[3].record()
        if record:
            torch.cuda.synchronize()
            for k, (i, j) in zip(ev, ((0, 1), (1, 2), (2, 3))):
                ev[k].append(marks[i].elapsed_time(marks[j]))
        return loss

    for _ in range(args.warmup):
        step(False)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    n0 = ops.launch_count()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        loss = step(True)
    t1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([t0.elapsed_time(t1) / args.steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.barrier()
    if rank == 0:
        mean = lambda xs: sum(xs) / max(1, len(xs))
        print(json.dumps({
            'metric': 'segments/sec stage-I contrastive training step (both encoders, fwd + bwd + optimizer)', 'value': B * S * world / (float(ms) / 1e3),
            'unit': 'segments/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': float(ms), 'dtype': 'bf16', 'data': 'synthetic',
            'scaling': 'weak', 'config': {'workload': f'segment_avclip.yaml training step, {B} x {S} segments/GPU', 'optimizer': 'torch.optim.AdamW'},
            'split_ms': {k: mean(v) for k, v in ev.items()}, 'gpu_launches_per_step': (ops.launch_count() - n0) / args.steps, 'loss': float(loss),
            'peak_mem_gb': torch.cuda.max_memory_allocated(dev) / 2 ** 30,
        }), flush=True)
    if world > 1:
        dist.destroy_process_group()


def avclip_forward_bench(args, rank, world, local, dev):
    """BASELINE config 3: segment_avclip.yaml contrastive FORWARD (AST + Motionformer, time-average pooled, similarity + CE), eval / no_grad,
    --batch x --segments segments per GPU (16 x 8 = 128 in the config)."""
    from synchformer_b200 import avclip, ops, synth
    B, S = args.batch, args.segments
    model = avclip.AVCLIP().to(dev).eval()
    sd = synth.synthetic_state_dict(1337, n_segments=S)
    model.v_encoder.load_state_dict({k[len('vfeat_extractor.'):]: v for k, v in sd.items() if k.startswith('vfeat_extractor.')})
    model.a_encoder.load_state_dict({k[len('afeat_extractor.'):]: v for k, v in sd.items() if k.startswith('afeat_extractor.')})
    g = torch.Generator(device=dev).manual_seed(rank)
    rgb = (torch.rand((B, S, 3, 16, 224, 224), device=dev, generator=g, dtype=torch.float16) - 0.5) / 0.5
    aud = torch.randn((B, S, 66, 128), device=dev, generator=g)
    with torch.no_grad():
        for _ in range(args.warmup):
            out = model(rgb, aud)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        n0 = ops.launch_count()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(args.steps):
            out = model(rgb, aud)
        t1.record()
        torch.cuda.synchronize()
    ms = torch.tensor([t0.elapsed_time(t1) / args.steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.barrier()
    if rank == 0:
        print(json.dumps({
            'metric': 'segments/sec stage-I contrastive forward (BASELINE config 3)', 'value': B * S * world / (float(ms) / 1e3), 'unit': 'segments/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': float(ms), 'dtype': 'bf16', 'data': 'synthetic', 'scaling': 'weak',
            'config': {'workload': f'segment_avclip.yaml contrastive forward, {B} x {S} segments/GPU'},
            'gpu_launches_per_step': (ops.launch_count() - n0) / args.steps, 'loss': float(out['losses']['segment_contrastive_loss']),
        }), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
