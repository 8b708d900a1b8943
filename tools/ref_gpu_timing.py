"""Reference-GPU denominator (SURVEY.md §8d "Reference GPU baseline"): the UNMODIFIED reference `Synchformer`
(model/sync_model.py:38-70), staged under the git-ignored baseline/_ref by tools/make_baseline_ref.py, timed on one B200 in eval /
no_grad through its own public call `model(vis, aud)`:
  (i)   fp16 autocast, as example.py:175 ships it          (ii) bf16 autocast          (iii) fp32 with TF32 matmuls
with `torch.backends.cudnn.benchmark = True` (scripts/train_sync.py:41), 3 warm-up + >= 10 timed iterations with CUDA events, at
BASELINE config 2's shape (S = 8, block_shape [114]) and the 5 s clip (S = 14), largest batch of the grid that fits.
A context number, NOT a bench value: it runs torch / cuBLAS / cuDNN kernels, none of this repo's.  Test infrastructure.

    SYNCHFORMER_REF=baseline/_ref python tools/ref_gpu_timing.py [out.json]
"""
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault('SYNCHFORMER_REF', os.path.join(REPO, 'baseline', '_ref'))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, 'tests', 'golden'))


def time_config(model, B, S, ctx_name, dev, iters=10, warm=3):
    ctxs = {'fp16_autocast': lambda: torch.autocast('cuda', dtype=torch.float16), 'bf16_autocast': lambda: torch.autocast('cuda', dtype=torch.bfloat16),
            'fp32_tf32': lambda: torch.autocast('cuda', enabled=False)}
    g = torch.Generator(device=dev).manual_seed(0)
    vis = (torch.rand(B, S, 16, 3, 224, 224, device=dev, generator=g) * 2 - 1)
    vis = vis.half() if ctx_name == 'fp16_autocast' else vis                  # RGBToHalfToZeroOne delivers fp16 video (sync.yaml:178-182)
    aud = torch.randn(B, S, 1, 128, 66, device=dev, generator=g)
    torch.cuda.reset_peak_memory_stats()
    with torch.no_grad(), ctxs[ctx_name]():
        for _ in range(warm):
            model(vis, aud)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            _, logits = model(vis, aud)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    return {'batch': B, 'segments': S, 'ms_per_step': ms, 'clips_per_s': B / (ms / 1e3), 'peak_mem_gb': torch.cuda.max_memory_allocated() / 2 ** 30,
            'logits_finite': bool(torch.isfinite(logits.float()).all())}


def main():
    out = os.path.abspath(sys.argv[1]) if len(sys.argv) > 1 else os.path.join(REPO, 'gpurun_out', 'reference_gpu.json')      # before import_reference() chdirs
    import _ref_import
    from synchformer_b200 import synth
    assert _ref_import.reference_available(), f"no reference at {os.environ['SYNCHFORMER_REF']} (run tools/make_baseline_ref.py where /root/reference exists)"
    dev = torch.device('cuda:0')
    torch.backends.cudnn.benchmark = True
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    res = {'what': 'UNMODIFIED reference model.sync_model.Synchformer.forward (baseline/_ref), eval, no_grad, cudnn.benchmark, CUDA events',
           'gpu': torch.cuda.get_device_name(0), 'torch': torch.__version__, 'runs': {}}
    for S in (8, 14):
        model = _ref_import.build_reference_model(S)
        model.load_state_dict(synth.synthetic_state_dict(1337, n_segments=S), strict=True)
        model = model.to(dev).eval()
        for name in ('fp16_autocast', 'bf16_autocast', 'fp32_tf32'):
            best = None
            for B in ((64, 32, 16, 8) if S == 8 else (32, 16, 8)):
                try:
                    r = time_config(model, B, S, name, dev, iters=10 if B <= 16 else 5)
                except torch.OutOfMemoryError:
                    torch.cuda.empty_cache()
                    continue
                except Exception as e:                     # noqa: BLE001  (record and go on: this is a measurement script)
                    r = {'batch': B, 'segments': S, 'error': repr(e)[:300]}
                res['runs'][f'S{S}_{name}_B{B}'] = r
                print(f'S={S} {name} B={B}: {r}', flush=True)
                if 'clips_per_s' in r and (best is None or r['clips_per_s'] > best['clips_per_s']):
                    best = r
                if 'clips_per_s' in r and B <= 16:
                    break                                  # largest batch that fits, plus one smaller point at most
            res[f'best_S{S}_{name}'] = best
        del model
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(out), exist_ok=True)
    with open(out, 'w') as f:
        json.dump(res, f, indent=1)
    print(json.dumps({k: v for k, v in res.items() if k.startswith('best_')}, indent=1))


if __name__ == '__main__':
    main()
