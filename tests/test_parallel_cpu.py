"""CPU: the N>1 host logic (segment partition, padded all-gathers, clip partition) with world_size-2 gloo processes."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from synchformer_b200 import parallel


def test_shard_range_covers_everything_once():
    for n in (1, 7, 14, 512, 513):
        for w in (1, 2, 3, 8):
            chunks = [parallel.shard_range(n, w, r) for r in range(w)]
            assert chunks[0][0] == 0 and chunks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(chunks, chunks[1:]))
            assert max(b - a for a, b in chunks) == -(-n // w)


def _fake_encode(all_feats):
    def f(s0, s1):
        return all_feats[s0:s1].clone()
    return f


def _fake_head(feats):     # (b, S, 14, D) -> (b, 3): any function of whole clips only
    return torch.stack([feats.sum((1, 2, 3)), feats[:, 0].mean((1, 2)), feats.amax((1, 2, 3))], dim=1)


def _worker(rank, world, port, B, S, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    feats = torch.randn(B * S, parallel.TOK_PER_SEG, 16, generator=g)
    out = parallel.sharded_forward(_fake_encode(feats), _fake_head, B, S)
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('B,S', [(4, 3), (1, 5), (3, 2)])
def test_sharded_forward_equals_single_process(B, S):
    g = torch.Generator().manual_seed(0)
    feats = torch.randn(B * S, parallel.TOK_PER_SEG, 16, generator=g)
    expect = _fake_head(feats.view(B, S, parallel.TOK_PER_SEG, 16))
    assert torch.equal(parallel.sharded_forward(_fake_encode(feats), _fake_head, B, S), expect)   # world 1 path
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() + B * 7 + S) % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, S, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(2):
        assert torch.equal(outs[r], expect), f'rank {r}'
