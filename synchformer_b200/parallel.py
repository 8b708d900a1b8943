"""Multi-GPU execution of the forward path: one process per GPU, `torch.distributed` (NCCL) for the plumbing.

The (clip x segment) encoder batch is embarrassingly parallel (segments are folded into the batch dimension in the
reference: motionformer.py:210, ast.py:162), and so is the sync transformer over clips.  One exchange sits between
them: rank r encodes a contiguous chunk of the flattened B*S segments and projects it, ONE all-gather
of the projected feature blocks gives every rank all segments, each rank runs the sync transformer on its own
clip range, and a second tiny all-gather returns the (B, n_cls) logits everywhere.  The reference itself has no
collective on this path (each rank owns whole clips); results are identical to the single-GPU forward because every
per-segment and per-clip computation is independent of its batch neighbours.

Nothing on this path synchronises the host: the class count comes from the module (no broadcast + `.item()`), the send /
receive buffers of both collectives are allocated once per (device, shape) and reused, and the `vproj` / `aproj` GEMM
epilogues write their fp32 rows straight into the send buffer of the feature all-gather - there is
no staging copy between the last kernel of the encoders and the collective.

Layout of one rank's send block (per = ceil(B*S / world) segments): [ per x 8 x 768 visual | per x 6 x 768 audio ], each
part a plain row-major matrix so that it can be a GEMM output; the receive buffer is `world` such blocks.

The host logic (partitioning, padding, gathers) takes the encoder / head as callables so that it is covered by
world_size-2 gloo tests on CPU (tests/test_parallel_cpu.py).
"""
from typing import Callable, Dict, Optional, Tuple

import torch
import torch.distributed as dist

TOK_PER_SEG = 14   # 8 visual + 6 audio tokens per segment after aggregation
V_TOK, A_TOK = 8, 6

_buffers: Dict[tuple, torch.Tensor] = {}


def _buffer(tag: str, shape: tuple, device, dtype) -> torch.Tensor:
    """Persistent buffer per (tag, shape, device, dtype): collectives and GEMM epilogues reuse the same storage every step (stable
    addresses: NCCL can keep them registered, CUDA graphs can capture them, the allocator is not touched on the hot path)."""
    key = (tag, tuple(shape), str(device), dtype)
    buf = _buffers.get(key)
    if buf is None:
        buf = torch.zeros(shape, device=device, dtype=dtype)       # zeros once: the padding rows of short ranks stay defined
        _buffers[key] = buf
    return buf


def shard_range(n_items: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous chunk [start, stop) of `n_items` for `rank`: ceil(n / world) items per rank, trailing ranks may be short or empty."""
    per = -(-n_items // world)
    start = min(rank * per, n_items)
    return start, min(start + per, n_items)


def all_gather_rows(local: torch.Tensor, n_total: int, world: int, group=None, tag: str = 'rows') -> torch.Tensor:
    """All-gather row blocks that were partitioned with `shard_range` (equal-size padded chunks, one collective).  The result is a view of a
    persistent receive buffer: valid until the next call with the same tag and shape."""
    per = -(-n_total // world)
    tail = tuple(local.shape[1:])
    send = _buffer(tag + '/send', (per,) + tail, local.device, local.dtype)
    if local.shape[0]:
        send[:local.shape[0]].copy_(local)
    out = _buffer(tag + '/recv', (world * per,) + tail, local.device, local.dtype)
    dist.all_gather_into_tensor(out, send, group=group)
    return out[:n_total]


def sharded_forward(encode_project: Callable[[int, int], torch.Tensor], sync_head: Callable[[torch.Tensor], torch.Tensor], B: int, S: int,
                    group=None, n_cls: Optional[int] = None) -> torch.Tensor:
    """Generic driver.
    encode_project(seg_start, seg_stop) -> (n_local, 14, D) features of flattened segments [seg_start, seg_stop): rows 0..7 projected
        visual tokens, rows 8..13 projected audio tokens.
    sync_head(feats (b, S, 14, D)) -> (b, n_cls) logits for whole clips.
    n_cls: width of the logits; required when some rank may own no clip (B < world), because such a rank cannot learn it from its own
        head call and nothing here asks another rank for it (that would be a host synchronisation).
    Returns logits (B, n_cls) on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    s0, s1 = shard_range(B * S, world, rank)
    local = encode_project(s0, s1)
    if world == 1:
        return sync_head(local.view(B, S, *local.shape[1:]))
    feats = all_gather_rows(local, B * S, world, group, tag='feats')          # the one exchange of the path
    feats = feats.view(B, S, *feats.shape[1:])
    c0, c1 = shard_range(B, world, rank)
    if c1 > c0:
        logits_local = sync_head(feats[c0:c1]).float()
        assert n_cls is None or logits_local.shape[-1] == n_cls, (logits_local.shape, n_cls)
        n_cls = logits_local.shape[-1]
    else:
        if n_cls is None:
            raise ValueError(f'rank {rank} owns no clip (B={B} < world={world}): pass n_cls')
        logits_local = torch.zeros((0, n_cls), device=feats.device, dtype=torch.float32)
    return all_gather_rows(logits_local, B, world, group, tag='logits')


def n_classes_of(model) -> int:
    """Width of the model's logits, read from the head's parameter (off_head: 21 offset classes; sync_head: 2)."""
    tr = model.transformer
    return int(getattr(tr, getattr(tr, '_HEAD', 'off_head')).weight.shape[0])


def synchformer_forward_sharded(model, vis_local: torch.Tensor, aud_local: torch.Tensor, B: int, S: int, group=None) -> torch.Tensor:
    """Offset-class logits (B, n_cls) for a GLOBAL batch of B clips x S segments, with this rank holding only its chunk of
    the flattened segments: vis_local (n_local, 16, 3, 224, 224), aud_local (n_local, 1, 128, 66) for segments
    shard_range(B*S, world, rank).  A view of a persistent buffer is returned (valid until the next call)."""
    D = 768
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    dev = vis_local.device
    s0, s1 = shard_range(B * S, world, rank)
    n = s1 - s0
    assert vis_local.shape[0] == n and aud_local.shape[0] == n, f'rank holds {vis_local.shape[0]} segments, expected {n}'
    per = -(-(B * S) // world)
    send = _buffer('sync/send', (per * TOK_PER_SEG, D), dev, torch.float32)
    send_v, send_a = send[:per * V_TOK], send[per * V_TOK:]
    if n > 0:
        vf = model.extract_vfeats(vis_local.unsqueeze(0))                  # (1, n, 8, 768)
        af = model.extract_afeats(aud_local.unsqueeze(0))                  # (1, n, 6, 768)
        model.project(vf, af, out_v=send_v[:n * V_TOK], out_a=send_a[:n * A_TOK])     # GEMM epilogues write into the send buffer
    if world == 1:
        v, a = send_v[:n * V_TOK].view(B, S * V_TOK, D), send_a[:n * A_TOK].view(B, S * A_TOK, D)
        return model.transformer(v, a)
    recv = _buffer('sync/recv', (world, per * TOK_PER_SEG, D), dev, torch.float32)
    dist.all_gather_into_tensor(recv.view(-1, D), send, group=group)                   # the one exchange of the path
    c0, c1 = shard_range(B, world, rank)
    n_cls = n_classes_of(model)
    if c1 > c0:
        # segments of clips [c0, c1) out of the per-rank blocks; (world * per, 8 | 6, 768) are strided views of the receive buffer
        seg0, seg1 = c0 * S, c1 * S
        v = _gather_segments(recv, seg0, seg1, per, 0, V_TOK, D).view(c1 - c0, S * V_TOK, D)
        a = _gather_segments(recv, seg0, seg1, per, per * V_TOK, A_TOK, D).view(c1 - c0, S * A_TOK, D)
        logits_local = model.transformer(v, a).float()
    else:
        logits_local = torch.zeros((0, n_cls), device=dev, dtype=torch.float32)
    return all_gather_rows(logits_local, B, world, group, tag='logits')


def _gather_segments(recv: torch.Tensor, seg0: int, seg1: int, per: int, part_offset: int, tok: int, D: int) -> torch.Tensor:
    """Rows of segments [seg0, seg1) of one modality out of the (world, per * 14, D) receive buffer -> contiguous ((seg1 - seg0) * tok, D).
    Segment s lives in block s // per at rows part_offset + (s % per) * tok; a clip range touches at most a few consecutive blocks."""
    pieces = []
    s = seg0
    while s < seg1:
        blk, off = divmod(s, per)
        cnt = min(per - off, seg1 - s)
        pieces.append(recv[blk, part_offset + off * tok: part_offset + (off + cnt) * tok])
        s += cnt
    return pieces[0] if len(pieces) == 1 else torch.cat(pieces, dim=0)
