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
    out = parallel.sharded_forward(_fake_encode(feats), _fake_head, B, S, n_cls=3)       # B < world: a rank without clips needs n_cls
    q.put((rank, out.clone()))
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


class _FakeTransformer(torch.nn.Module):
    _HEAD = 'off_head'

    def __init__(self):
        super().__init__()
        self.off_head = torch.nn.Linear(768, 5)

    def forward(self, v, a):        # any function of whole clips: (b, 8S, 768), (b, 6S, 768) -> (b, 5)
        w = torch.arange(1, v.shape[1] + 1, dtype=v.dtype).view(1, -1, 1)
        u = torch.arange(1, a.shape[1] + 1, dtype=a.dtype).view(1, -1, 1)
        return torch.stack([(v * w).sum((1, 2)), (a * u).sum((1, 2)), v[:, 0, :3].sum(1), a[:, -1, :3].sum(1), v.amax((1, 2)) + a.amin((1, 2))], dim=1)


class _FakeModel:
    """stands in for Synchformer in the host logic of synchformer_forward_sharded: per-segment encoders, projections that write into the
    caller's buffers (the all-gather send buffer), a per-clip head"""

    def __init__(self):
        self.transformer = _FakeTransformer()

    def extract_vfeats(self, vis):          # (1, n, 16, 3, 224, 224) stand-in: (1, n, 4) -> (1, n, 8, 768)
        return vis.view(1, -1, 1, 4).mean(-1, keepdim=True) * torch.linspace(0.5, 1.5, 8 * 768).view(1, 1, 8, 768)

    def extract_afeats(self, aud):          # (1, n, 3) -> (1, n, 6, 768)
        return aud.view(1, -1, 1, 3).sum(-1, keepdim=True) * torch.linspace(-1.0, 1.0, 6 * 768).view(1, 1, 6, 768)

    def project(self, vf, af, out_v=None, out_a=None):
        out_v.copy_(vf.reshape(-1, 768) * 2.0)
        out_a.copy_(af.reshape(-1, 768) - 1.0)
        return out_v, out_a


def _model_worker(rank, world, port, B, S, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    g = torch.Generator().manual_seed(1)
    vis, aud = torch.randn(B * S, 4, generator=g), torch.randn(B * S, 3, generator=g)
    s0, s1 = parallel.shard_range(B * S, world, rank)
    outs = []
    for _ in range(2):                     # second call reuses the persistent buffers
        outs.append(parallel.synchformer_forward_sharded(_FakeModel(), vis[s0:s1], aud[s0:s1], B, S).clone())
    assert torch.equal(outs[0], outs[1])
    q.put((rank, outs[0]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('B,S', [(3, 3), (1, 5), (2, 4)])
def test_synchformer_forward_sharded_host_logic(B, S):
    """The product entry point with a stand-in model: send-buffer layout [visual | audio], projections written into it, block-wise
    re-assembly of a rank's clips from the receive buffer, no n_cls broadcast - equal to the single-process result on every rank."""
    g = torch.Generator().manual_seed(1)
    vis, aud = torch.randn(B * S, 4, generator=g), torch.randn(B * S, 3, generator=g)
    expect = parallel.synchformer_forward_sharded(_FakeModel(), vis, aud, B, S).clone()       # world 1 path
    assert expect.shape == (B, 5)
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 31500 + (os.getpid() + B * 11 + S) % 2000
    procs = [ctx.Process(target=_model_worker, args=(r, 2, port, B, S, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = dict(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(2):
        assert torch.equal(outs[r], expect), f'rank {r}'
