"""Multi-GPU execution of the forward path: one process per GPU, `torch.distributed` (NCCL) for the plumbing.

The (clip x segment) encoder batch is embarrassingly parallel (segments are folded into the batch dimension in the
reference: motionformer.py:210, ast.py:162), and so is the sync transformer over clips.  One exchange sits between
them: rank r encodes a contiguous chunk of the flattened B*S segments and projects it, ONE all-gather
of the (n_local, 14, 768) feature blocks gives every rank all segments, each rank runs the sync transformer on its own
clip range, and a second tiny all-gather returns the (B, n_cls) logits everywhere.  The reference itself has no
collective on this path (each rank owns whole clips); results are identical to the single-GPU forward because every
per-segment and per-clip computation is independent of its batch neighbours.

The host logic (partitioning, padding, gathers) takes the encoder / head as callables so that it is covered by
world_size-2 gloo tests on CPU (tests/test_parallel_cpu.py).
"""
from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist

TOK_PER_SEG = 14   # 8 visual + 6 audio tokens per segment after aggregation


def shard_range(n_items: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous chunk [start, stop) of `n_items` for `rank`: ceil(n / world) items per rank, trailing ranks may be short or empty."""
    per = -(-n_items // world)
    start = min(rank * per, n_items)
    return start, min(start + per, n_items)


def all_gather_rows(local: torch.Tensor, n_total: int, world: int, group=None) -> torch.Tensor:
    """All-gather row blocks that were partitioned with `shard_range` (equal-size padded chunks, one collective)."""
    per = -(-n_total // world)
    pad = torch.zeros((per,) + tuple(local.shape[1:]), device=local.device, dtype=local.dtype)
    pad[:local.shape[0]] = local
    out = torch.empty((world * per,) + tuple(local.shape[1:]), device=local.device, dtype=local.dtype)
    dist.all_gather_into_tensor(out, pad, group=group)
    return out[:n_total]


def sharded_forward(encode_project: Callable[[int, int], torch.Tensor], sync_head: Callable[[torch.Tensor], torch.Tensor], B: int, S: int,
                    group=None) -> torch.Tensor:
    """Generic driver.
    encode_project(seg_start, seg_stop) -> (n_local, 14, D) features of flattened segments [seg_start, seg_stop): rows 0..7 projected
        visual tokens, rows 8..13 projected audio tokens.
    sync_head(feats (b, S, 14, D)) -> (b, n_cls) logits for whole clips.
    Returns logits (B, n_cls) on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    s0, s1 = shard_range(B * S, world, rank)
    local = encode_project(s0, s1)
    if world == 1:
        return sync_head(local.view(B, S, *local.shape[1:]))
    feats = all_gather_rows(local, B * S, world, group)                       # the one exchange of the path
    feats = feats.view(B, S, *feats.shape[1:])
    c0, c1 = shard_range(B, world, rank)
    if c1 > c0:
        logits_local = sync_head(feats[c0:c1])
        n_cls = logits_local.shape[-1]
    else:
        logits_local, n_cls = None, None
    # ranks without clips still need n_cls for the gather: broadcast it from rank 0 (which always owns clip 0)
    n_t = torch.tensor([n_cls if n_cls is not None else 0], device=feats.device, dtype=torch.int64)
    dist.broadcast(n_t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    n_cls = int(n_t.item())
    if logits_local is None:
        logits_local = torch.zeros((0, n_cls), device=feats.device, dtype=torch.float32)
    return all_gather_rows(logits_local.float(), B, world, group)


def synchformer_forward_sharded(model, vis_local: torch.Tensor, aud_local: torch.Tensor, B: int, S: int, group=None) -> torch.Tensor:
    """Offset-class logits (B, n_cls) for a GLOBAL batch of B clips x S segments, with this rank holding only its chunk of
    the flattened segments: vis_local (n_local, 16, 3, 224, 224), aud_local (n_local, 1, 128, 66) for segments
    shard_range(B*S, world, rank)."""
    D = 768

    def encode_project(s0: int, s1: int) -> torch.Tensor:
        n = s1 - s0
        assert vis_local.shape[0] == n and aud_local.shape[0] == n, f'rank holds {vis_local.shape[0]} segments, expected {n}'
        out = torch.empty((n, TOK_PER_SEG, D), device=vis_local.device, dtype=torch.float32)
        if n == 0:
            return out
        vf = model.extract_vfeats(vis_local.unsqueeze(0))                  # (1, n, 8, 768)
        af = model.extract_afeats(aud_local.unsqueeze(0))                  # (1, n, 6, 768)
        v, a = model.project(vf, af)                                       # (1, 8n, 768), (1, 6n, 768)
        out[:, :8] = v.view(n, 8, D)
        out[:, 8:] = a.view(n, 6, D)
        return out

    def sync_head(feats: torch.Tensor) -> torch.Tensor:
        b = feats.shape[0]
        v = feats[:, :, :8].reshape(b, S * 8, D)
        a = feats[:, :, 8:].reshape(b, S * 6, D)
        return model.transformer(v, a)

    return sharded_forward(encode_project, sync_head, B, S, group)
