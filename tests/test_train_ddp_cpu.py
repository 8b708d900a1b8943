"""CPU: the training step under DistributedDataParallel with world_size-2 gloo processes (BASELINE config 4's host logic,
scripts/train_utils.py:205-210): the autograd Functions of synchformer_b200/train.py feed DDP's gradient hooks, each rank runs its
half of the batch, and the all-reduced gradients equal the full-batch oracle gradients.  Kernels are the CPU stand-ins of
tests/fake_ops.py (fp32), so this checks the plumbing, not the arithmetic of the kernels."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
B, S = 4, 2


def _inputs():
    g = torch.Generator().manual_seed(11)
    vf = torch.randn((B, S, 8, 768), generator=g) * 0.5
    af = torch.randn((B, S, 6, 768), generator=g) * 0.5
    return vf, af, torch.tensor([1, 5, 20, 9])


def _worker(rank, world, port, q):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    import fake_ops
    from synchformer_b200 import model as M, synth, train
    dist.init_process_group('gloo', rank=rank, world_size=world)
    mpatch = pytest.MonkeyPatch()
    fake_ops.install(mpatch, round_bf16=False)
    cfg = M.sync_yaml_model_config(S)
    for k in ('embd_pdrop', 'resid_pdrop', 'attn_pdrop'):
        cfg['transformer']['params'][k] = 0.0                     # masks depend on the per-rank batch layout; p = 0 makes ranks comparable
    cfg = {k: {kk: vv for kk, vv in v.items() if kk != 'is_trainable'} for k, v in cfg.items()}
    model = M.Synchformer(**cfg)
    model.load_state_dict(synth.synthetic_state_dict(1337, n_segments=S), strict=True)
    model.train()
    for ext in (model.vfeat_extractor, model.afeat_extractor):   # train_utils.py:199-204
        ext.requires_grad_(False)
        ext.eval()

    class _SyncOnly(torch.nn.Module):                             # DDP wraps a module whose forward is the sync-module half of Synchformer.forward
        def __init__(self, m):
            super().__init__()
            self.m = m

        def forward(self, vf, af, targets):
            v, a = self.m.project(vf, af)
            logits = self.m.transformer(v, a)
            return self.m.compute_loss(logits, targets), logits

    ddp = torch.nn.parallel.DistributedDataParallel(_SyncOnly(model))
    vf, af, targets = _inputs()
    lo, hi = rank * B // world, (rank + 1) * B // world
    with torch.enable_grad():
        loss, _ = ddp(vf[lo:hi], af[lo:hi], targets[lo:hi])
        loss.backward()
    grads = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
    q.put((rank, {n: g.numpy() for n, g in grads.items()}))
    dist.barrier()
    dist.destroy_process_group()
    mpatch.undo()


def test_ddp_gradients_equal_full_batch_oracle():
    sys.path.insert(0, HERE)
    from oracle import synchformer_oracle as O
    from synchformer_b200 import synth
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = dict(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    vf, af, targets = _inputs()
    sd = synth.synthetic_state_dict(1337, n_segments=S)
    _, _, ref = O.sync_train_grads(sd, vf, af, targets, None)       # mean CE over the full batch == mean of the two half-batch means
    assert set(outs[0]) == set(ref) and len(ref) == 63
    for n, r in ref.items():
        assert np.array_equal(outs[0][n], outs[1][n]), n            # all-reduced: identical on both ranks
        err, scale = float(np.linalg.norm(outs[0][n] - r.numpy())), float(r.norm())
        assert err <= 2e-4 * scale + 1e-6, (n, err, scale)
