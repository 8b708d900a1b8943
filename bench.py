#!/usr/bin/env python
"""Throughput benchmark of the Synchformer forward path on B200 (BASELINE.json metric: clips/sec of offset inference).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch 64] [--segments 8] [--impl b200|reference]

A step = one pass of the hot path over one batch of synthetic clips (per GPU): raw 16 kHz waveform -> GPU mel front-end,
fp16 RGB segments -> Motionformer, AST, CLS aggregators, projections, [all-gather of segment features when N > 1],
sync transformer -> (B, 21) logits.  Workload at N = 1 is BASELINE.json configs[1]: sync.yaml inference, batch 64,
8 segments / clip, bf16.  Scaling is weak: every rank processes `--batch` clips, `value` = all clips / max-over-ranks time.

Rank 0 prints ONE JSON line (see the keys in main()).  `--impl reference` times the reference's CPU path on the host cores, rank 0 only: the UNMODIFIED
reference staged under the git-ignored baseline/_ref (tools/make_baseline_ref.py; kind "reference") or, where that is absent, the CPU
oracle restatement (`oracle/`, kind "port").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

GFLOP_PER_SEGMENT = 408.260            # BASELINE.md §3 (canonical reference op graph, 2 x MAC)


def flops_per_clip(S: int) -> float:
    T = 2 + 14 * S
    return (S * GFLOP_PER_SEGMENT * 1e9 + 2 * (14 * S) * 768 ** 2 + 3 * (24 * T * 768 ** 2 + 8 * 4 * T * T * 96) + 2 * 768 * 21)


def measured_peaks():
    p = os.path.join(REPO, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(sustained=float(d.get('bf16_tflops_sustained', 1395.6)), burst=float(d.get('bf16_tflops', 1696.6)),
                    hbm=float(d.get('hbm_gbs', 6445.0)), source='measured (MEASURED_PEAKS.json)')
    return dict(sustained=1400.0, burst=1590.0, hbm=6650.0, source='fallback (B200_PROFILING.md)')


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-i', str(self.index), '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace('.', '').isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith('active') for r in self.rows)]
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None, 'reasons': reasons, 'samples': len(sm)}


def _staged_reference_root():
    """The unmodified reference staged by tools/make_baseline_ref.py under the git-ignored baseline/_ref (it travels with the gpurun
    snapshot; /root/reference does not exist on the GPU box)."""
    root = os.path.join(REPO, 'baseline', '_ref')
    return root if os.path.isdir(os.path.join(root, 'model')) else None


def cpu_reference_clips_per_sec(S: int, n_clips: int, steps: int, warmup: int):
    """The reference's own CPU path on the host cores, fp32, all threads, bounded sample.  kind 'reference': the UNMODIFIED reference
    (baseline/_ref: its transform tail dataset/transforms.py:815-871 + model.sync_model.Synchformer.forward) when it is staged; else kind
    'port': the torch-CPU oracle restatement.  Returns (clips/s, seconds per step, kind, description)."""
    from synchformer_b200 import synth
    torch.set_num_threads(os.cpu_count() or 1)
    sd = synth.synthetic_state_dict(1337, n_segments=S)
    vis = synth.synthetic_video(n_clips, S, 0)
    wave = synth.synthetic_waveform(n_clips, S, 0)
    root = _staged_reference_root()
    cwd = os.getcwd()
    if root is not None:
        os.environ['SYNCHFORMER_REF'] = root
        sys.path.insert(0, os.path.join(REPO, 'tests', 'golden'))
        import importlib
        import _ref_import
        _ref_import.REF_ROOT = root
        model = _ref_import.build_reference_model(S)
        model.load_state_dict(sd, strict=True)
        T = importlib.import_module('dataset.transforms')
        chain = [T.AudioMelSpectrogram(sample_rate=16000, win_length=400, hop_length=160, n_fft=1024, n_mels=128), T.AudioLog(),
                 T.PadOrTruncate(max_spec_t=66), T.AudioNormalizeAST(mean=-4.2677393, std=4.5689974)]          # configs/sync.yaml:183-197

        def step():
            auds = []
            for b in range(n_clips):
                item = {'audio': wave[b], 'meta': {'audio': {}}}
                for t in chain:
                    item = t(item)
                auds.append(item['audio'])
            return model(vis, torch.stack(auds).unsqueeze(2))[1]                                             # (B, S, 1, F, T)
        kind, what = 'reference', 'UNMODIFIED reference (baseline/_ref): torchaudio mel transform tail + model.sync_model.Synchformer.forward'
    else:
        from oracle import synchformer_oracle as O

        def step():
            return O.forward(sd, vis, O.mel_frontend(wave).float().unsqueeze(2))[1]
        kind, what = 'port', 'torch CPU oracle port of the reference forward (mel + encoders + sync)'
    times = []
    try:
        with torch.no_grad():
            for it in range(warmup + steps):
                t0 = time.perf_counter()
                step()
                if it >= warmup:
                    times.append(time.perf_counter() - t0)
    finally:
        os.chdir(cwd)                       # importing the reference changes the working directory
    sec = sum(times) / len(times)
    return n_clips / sec, sec, kind, what


def run_reference(args, rank: int):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_clips = 1
    value, sec, kind, what = cpu_reference_clips_per_sec(args.segments, n_clips, max(1, min(args.steps, 2)), min(args.warmup, 1))
    sample = f'{n_clips} clip x {args.segments} segments per step, fp32, {what}'
    print(json.dumps({
        'impl': 'reference', 'metric': 'clips/sec offset inference (5s-style clip: S x 0.64 s segments, 224p RGB + 16 kHz)', 'value': value,
        'unit': 'clips/s', 'n_gpus': args.gpus, 'steps': max(1, min(args.steps, 2)), 'warmup': min(args.warmup, 1), 'ms_per_step': sec * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f'sync.yaml inference, batch={args.batch}, {args.segments} segments/clip (CPU sample: {n_clips} clip/step)'},
        'cpu_baseline': {'value': value, 'unit': 'clips/s', 'cores': cores, 'kind': kind, 'sample': sample},
        'e2e': {'value': value, 'unit': 'clips/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }), flush=True)


class GemmTimer:
    """CUDA-event timing of every launch of the hot kernels inside the timed region, on the launching stream: the GEMM (roofline.achieved:
    algorithmic FLOPs / time) and the HBM-bound kernel classes around it (attention by shape class, LayerNorm: algorithmic bytes / time)."""

    def __init__(self, ops):
        self.ops, self.records, self.other = ops, [], {}
        self.orig = {n: getattr(ops, n) for n in ('gemm', 'attention', 'layernorm', 'rowstats_cast')}

    def _timed(self, fn, key_work):
        def wrapper(*a, **kw):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            r = fn(*a, **kw)
            e.record()
            key, work = key_work(*a, **kw)
            if key == 'gemm':
                self.records.append((work, s, e))
            else:
                self.other.setdefault(key, []).append((work, s, e))
            return r
        return wrapper

    def __enter__(self):
        def gemm_work(a, w, bias, out=None, **kw):
            return 'gemm', 2.0 * a.shape[0] * w.shape[0] * w.shape[1]

        def attn_work(q, k, v, out, *, n_outer, n_inner, n_heads, head_dim, Lq, Lk, k_prefix=None, **kw):
            cls = 'attn_space_196x197' if Lq == 196 else 'attn_time_8x9' if (Lq == 8 and Lk == 8) else 'attn_cls_row' if Lq == 1 else f'attn_{Lq}x{Lk}_hd{head_dim}'
            rows = n_outer * n_inner * n_heads * (2 * Lq + 2 * Lk) + (2 * n_outer * n_heads if k_prefix is not None else 0)
            return cls, float(rows * head_dim * 2)                     # q + out rows, k + v rows (bf16), every row once

        def ln_work(x, *a, rows=None, out_f32=False, **kw):
            r = x.shape[0] if rows is None else rows
            return 'layernorm', float(r * 768 * (4 + (4 if out_f32 else 2)))

        def rs_work(x, *a, **kw):
            return 'layernorm', float(x.shape[0] * 768 * 6)

        for name, kw in (('gemm', gemm_work), ('attention', attn_work), ('layernorm', ln_work), ('rowstats_cast', rs_work)):
            setattr(self.ops, name, self._timed(self.orig[name], kw))
        return self

    def __exit__(self, *exc):
        for name, fn in self.orig.items():
            setattr(self.ops, name, fn)

    def summary(self):
        flops = sum(f for f, _, _ in self.records)
        ms = sum(s.elapsed_time(e) for _, s, e in self.records)
        return flops, ms, len(self.records)

    def hbm_kernels(self, steps: int, hbm_peak_gbs: float):
        """per kernel class: launches / step, ms / step, achieved GB/s of ALGORITHMIC bytes (every operand row once), fraction of the measured copy peak"""
        out = {}
        for key, recs in sorted(self.other.items()):
            ms = sum(s.elapsed_time(e) for _, s, e in recs)
            nbytes = sum(b for b, _, _ in recs)
            gbs = nbytes / (ms / 1e3) / 1e9 if ms > 0 else None
            out[key] = {'launches_per_step': len(recs) / steps, 'ms_per_step': ms / steps, 'achieved_gbs': gbs, 'frac_of_hbm_peak': gbs / hbm_peak_gbs if gbs else None}
        return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--batch', type=int, default=64, help='clips per GPU per step')
    ap.add_argument('--segments', type=int, default=8)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup

    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        run_reference(args, rank)
        return

    import torch.distributed as dist
    from synchformer_b200 import model as M, ops, parallel, synth
    if not torch.cuda.is_available():
        raise SystemExit('bench.py --impl b200 needs a CUDA device (there is no CPU path)')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    ops.device_check()

    B, S = args.batch, args.segments                    # per-rank clips; global batch = B * world (weak scaling)
    model = M.build_synchformer(n_segments=S, state_dict=synth.synthetic_state_dict(1337, n_segments=S), device=dev)
    g = torch.Generator(device=dev).manual_seed(rank)
    # inputs resident in HBM: fp16 video as RGBToHalfToZeroOne delivers it (2.47 GB per rank at B=64,S=8 >> 126 MB L2), raw waveform
    vis = (torch.rand(B, S, 16, 3, 224, 224, device=dev, generator=g) * 2 - 1).half()
    t = torch.arange(S * 5120 + 5120, device=dev, dtype=torch.float32) / 16000.0
    wave = torch.stack([torch.sin(2 * torch.pi * 440.0 * 2 ** (b / 12.0) * t).unfold(0, 10240, 5120)[:S] for b in range(B)]).contiguous()
    # host copies for the end-to-end leg (pinned)
    vis_h = vis.cpu().pin_memory()
    wave_h = wave.cpu().pin_memory()

    def step(v, w):
        with torch.no_grad():
            mel = ops.mel_frontend(w).unsqueeze(2)                      # (B, S, 1, 128, 66)
            if world == 1:
                return model(v, mel)[1]
            # this rank's clips are its contiguous chunk of the global flattened segment list
            return parallel.synchformer_forward_sharded(model, v.view(B * S, 16, 3, 224, 224), mel.view(B * S, 1, 128, 66), B * world, S)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        logits = step(vis, wave)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = ops.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with GemmTimer(ops) as gt:
        barrier()
        ev0.record()
        for _ in range(args.steps):
            logits = step(vis, wave)
        ev1.record()
        barrier()
    ms = ev0.elapsed_time(ev1)
    launches = ops.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    gemm_flops, gemm_ms, gemm_n = gt.summary()

    # end-to-end: the public call (Synchformer.forward) fed from pinned HOST buffers, logits read back, all inside the timed region.
    # Two device input buffers and a copy stream: the H2D copy of step i+1 runs while step i computes (what a prefetching
    # DataLoader + non_blocking .to() gives the reference harness); every step still pays its own copy and its own read-back.
    bufs = [(vis, wave), (torch.empty_like(vis), torch.empty_like(wave))]
    copy_stream = torch.cuda.Stream(device=dev)
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]

    def prefetch(b):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[b])
            bufs[b][0].copy_(vis_h, non_blocking=True)
            bufs[b][1].copy_(wave_h, non_blocking=True)
            ready[b].record(copy_stream)

    def e2e_run(n):
        for b in range(2):
            consumed[b].record()
        prefetch(0)
        out = None
        for i in range(n):
            b = i & 1
            if i + 1 < n:
                prefetch(b ^ 1)
            torch.cuda.current_stream().wait_event(ready[b])
            lg = step(*bufs[b])
            consumed[b].record()
            out = lg.float().cpu()                               # device -> host read of the step's result
        return out

    e2e_run(2)
    barrier()
    e2e_steps = max(3, min(args.steps, 5))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out_h = e2e_run(e2e_steps)
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)

    tm = torch.tensor([ms, e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    ms, e2e_ms = float(tm[0]), float(tm[1])

    if rank == 0:
        peaks = measured_peaks()
        value = B * world * args.steps / (ms / 1e3)
        e2e_value = B * world * e2e_steps / (e2e_ms / 1e3)
        achieved = gemm_flops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else None
        traffic, traffic_note = None, None
        tp = os.path.join(REPO, 'profiles', 'r2_gemm_traffic.json')
        if os.path.exists(tp):                      # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch from the committed ncu capture
            tj = json.load(open(tp))
            traffic, traffic_note = tj['dram_bytes'], f"{tj['launch']}; algorithmic bytes {tj['algorithmic_bytes']}; {tj['source']}"
        result = {
            'metric': 'clips/sec offset inference (5s-style clip: S x 0.64 s segments, 224p RGB + 16 kHz)', 'value': value, 'unit': 'clips/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
            'config': {'workload': f'sync.yaml inference, batch={B} clips/GPU, {S} segments/clip, bf16 GEMMs (fp32 accumulate, fp32 residual stream)',
                       'global_batch': B * world, 'segments': S, 'parallelism': f'dp{world} (segment-sharded encoders + 1 all-gather)' if world > 1 else 'single GPU',
                       'l2_policy': 'inputs larger than L2 (2.47 GB fp16 video per rank); activations stream through HBM',
                       'weights': 'synthetic_state_dict(seed 1337), random-init-like, non-zero patch embedding'},
            'roofline': {'bound': 'tensor', 'achieved': achieved, 'peak': peaks['sustained'], 'unit': 'TFLOP/s',
                         'frac': (achieved / peaks['sustained']) if achieved else None, 'traffic': traffic, 'traffic_launch': traffic_note,
                         'achieved_def': 'sum over the GEMM launches of the timed steps of 2*M*N*K / sum of their CUDA-event durations',
                         'kernel': 'gemm_bf16_tcgen05_kernel', 'launches_timed': gemm_n, 'gemm_ms_per_step': gemm_ms / args.steps,
                         'peak_source': peaks['source'] + ', sustained bf16 (kernel timed inside a long step)',
                         'step_frac_canonical': value * flops_per_clip(S) / world / (peaks['sustained'] * 1e12),
                         'step_frac_note': 'whole step (all kernels) against the canonical reference FLOP count; `frac` above is the dominant kernel alone',
                         'flops_per_clip_canonical': flops_per_clip(S),
                         # the MHSA / LayerNorm kernels are HBM-bound (AI of the 196 x 197 space attention ~ 99 FLOP/B against a ridge of 223),
                         # so their roofline is the measured copy bandwidth, not the tensor pipe
                         'hbm_bound_kernels': gt.hbm_kernels(args.steps, peaks['hbm']), 'hbm_peak_gbs': peaks['hbm']},
            'clocks': clocks,
            'e2e': {'value': e2e_value, 'unit': 'clips/s', 'h2d_bytes_per_step': vis_h.numel() * vis_h.element_size() + wave_h.numel() * 4,
                    'd2h_bytes_per_step': out_h.numel() * 4, 'steps': e2e_steps},
            'gpu_launches': launches,
            'logits_checksum': float(logits.float().abs().sum()),
        }
        if world == 1 and not args.no_cpu_baseline:
            cv, csec, ckind, cwhat = cpu_reference_clips_per_sec(S, 1, 1, 0)
            result['cpu_baseline'] = {'value': cv, 'unit': 'clips/s', 'cores': os.cpu_count() or 1, 'kind': ckind,
                                      'sample': f'1 clip x {S} segments, fp32, {cwhat}, {csec:.1f} s'}
        print(json.dumps(result), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
