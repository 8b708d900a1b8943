"""Deterministic synthetic weights and inputs (there is no network for checkpoints / datasets).

`synthetic_state_dict` fills the reference's state-dict schema with "trained-like" values: every tensor is
drawn from its own `torch.Generator` seeded by (seed, crc32(name)), so the result does not depend on
construction order and is reproducible on any host with the same torch build.  Unlike the reference's random
init it gives the Conv3d patch embedding non-zero weights (video_model_builder.py:61 zeroes them, which makes
visual features input-independent — SURVEY.md §0 trap 1) and jitters LayerNorm gains/biases, so that every
kernel on the path influences the outputs that the parity tests compare.

Inputs follow BASELINE.md §4: video U(-1, 1) shaped (B, S, 16, 3, 224, 224); audio = 16 kHz sine cut into S
windows of 10 240 samples at stride 5 120.
"""
import math
import zlib
from typing import Dict

import torch

from .schema import state_dict_schema

SEG_SAMPLES = 10240           # 0.64 s @ 16 kHz  (configs/sync.yaml:93-101)
SEG_STRIDE = 5120


def _gen(seed: int, name: str) -> torch.Generator:
    g = torch.Generator(device='cpu')
    g.manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 63 - 1))
    return g


def synthetic_state_dict(seed: int = 1337, n_segments: int = 14, n_classes: int = 21, head: str = 'off_head',
                         w_std: float = 0.03) -> Dict[str, torch.Tensor]:
    sd = {}
    for name, shape in state_dict_schema(n_segments, n_classes, head).items():
        g = _gen(seed, name)
        leaf = name.rsplit('.', 1)[-1]
        is_ln = any(t in name for t in ('norm', 'lnorm', 'ln1', 'ln2', 'ln_f', 'layernorm'))
        if is_ln and leaf == 'weight':
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif is_ln and leaf == 'bias':
            t = 0.05 * torch.randn(shape, generator=g)
        elif leaf in ('bias', 'in_proj_bias'):
            t = 0.02 * torch.randn(shape, generator=g)
        elif leaf in ('weight', 'in_proj_weight'):
            fan_in = math.prod(shape[1:])
            std = w_std if fan_in <= 1536 else w_std * math.sqrt(768.0 / fan_in)
            t = std * torch.randn(shape, generator=g)
        else:   # tokens and positional tables
            t = 0.2 * torch.randn(shape, generator=g)
        sd[name] = t.float()
    return sd


def synthetic_video(B: int, S: int, seed: int = 0, dtype=torch.float32) -> torch.Tensor:
    """(B, S, 16, 3, 224, 224) in [-1, 1]: uint8-like frames -> /255 -> (x - .5) / .5
    (dataset/transforms.py:647-669).  A smooth per-clip pattern plus noise, so frames differ in space and time."""
    g = _gen(seed, f'video{B}x{S}')
    u8 = torch.randint(0, 256, (B, S, 16, 3, 224, 224), generator=g, dtype=torch.int32)
    x = (u8.float() / 255.0 - 0.5) / 0.5
    return x.to(dtype)


def synthetic_waveform(B: int, S: int, seed: int = 0, freq_hz: float = 440.0) -> torch.Tensor:
    """(B, S, 10240) float32: per-clip sine (clip b is detuned by b semitones) cut into overlapping windows."""
    n = SEG_STRIDE * (S - 1) + SEG_SAMPLES
    t = torch.arange(n, dtype=torch.float64) / 16000.0
    clips = []
    for b in range(B):
        f = freq_hz * 2.0 ** (b / 12.0)
        w = torch.sin(2.0 * math.pi * f * t) + 0.25 * torch.sin(2.0 * math.pi * 3.1 * f * t + 0.5)
        clips.append(w.unfold(0, SEG_SAMPLES, SEG_STRIDE))
    return torch.stack(clips).float().contiguous()
