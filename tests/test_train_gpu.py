"""GPU parity of the training step of the synchronisation module: every new kernel against the CPU restatement of
its contract (tests/fake_ops.py, oracle/philox.py), and the whole step (forward with dropout, backward, 63 gradient tensors) against
torch autograd on the fp32 oracle with the same dropout multipliers.

These tests gate: they were seen green on a B200 (GPUTEST_r01: 42 / 42 of the training tests passed) and carry no xfail marker.
"""
import copy
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import philox
from oracle import synchformer_oracle as O
from synchformer_b200 import model as M, ops, synth, train

import fake_ops
import train_gates

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
D = 768


@pytest.fixture(autouse=True)
def _grad_enabled():
    """other test modules import the reference, which switches autograd off process-wide"""
    with torch.enable_grad():
        yield


def _bf(x):
    return x.to(torch.bfloat16)


# ---- elementwise / reduction kernels -----------------------------------------------------------------------------------------
@pytest.mark.parametrize('p', [0.0, 0.1, 0.5])
def test_dropout_mask_is_the_philox_oracle_bit_for_bit(cuda_device, p):
    n_rows, seed, site = 333, 0x1234_5678_9ABC_DEF, 7
    ones = torch.ones((n_rows, D), device=cuda_device)
    got = ops.dropout(ones, p, seed, site).cpu().numpy()
    want = philox.dropout_multiplier((n_rows, D), p, seed, site)
    assert np.array_equal(got, want)
    x = torch.randn((n_rows, D), device=cuda_device)
    res = torch.randn((n_rows, D), device=cuda_device)
    y = ops.dropout(x, p, seed, site, residual=res)
    assert torch.equal(y.cpu(), res.cpu() + x.cpu() * torch.from_numpy(want))
    yb = ops.dropout(x, p, seed, site, out_bf16=True)
    assert torch.equal(yb.cpu(), _bf(x.cpu() * torch.from_numpy(want)))
    ops.dropout(x, p, seed, site, out=x)                                                   # in place
    assert torch.equal(x.cpu(), (y - res).cpu()) or torch.allclose(x.cpu(), (y - res).cpu(), atol=1e-6)


@pytest.mark.parametrize('shape', [(594, 768), (60, 3072), (7, 2304), (1000, 40)])
def test_transpose_bf16_pads_with_zeros(cuda_device, shape):
    x = _bf(torch.randn(shape, device=cuda_device))
    t = ops.transpose_bf16(x)
    R, C = shape
    assert t.shape == (C, (R + 7) // 8 * 8)
    assert torch.equal(t[:, :R].cpu(), x.cpu().t())
    assert float(t[:, R:].float().abs().sum()) == 0.0
    big = _bf(torch.randn((R, C + 16), device=cuda_device))                                # strided input (row stride > C)
    assert torch.equal(ops.transpose_bf16(big[:, :C])[:, :R].cpu(), big[:, :C].cpu().t())


@pytest.mark.parametrize('M_,N_,dtype', [(594, 768, torch.bfloat16), (6336, 3072, torch.bfloat16), (3, 152064, torch.float32),
                                         (5000, 2304, torch.float32), (1, 768, torch.float32)])
def test_colsum(cuda_device, M_, N_, dtype):
    x = torch.randn((M_, N_), device=cuda_device).to(dtype)
    got = ops.colsum(x).cpu().double()
    want = x.cpu().double().sum(0)
    assert (got - want).abs().max() <= 1e-5 * x.cpu().double().abs().sum(0).max() + 1e-6
    assert torch.equal(ops.colsum(x).cpu().double(), got)                                  # deterministic


def test_gelu_forward_and_backward(cuda_device):
    x = _bf(torch.randn((1024, 3072), device=cuda_device) * 2.0)
    dy = _bf(torch.randn((1024, 3072), device=cuda_device))
    y = ops.gelu_fwd(x).cpu().float()
    want = F.gelu(x.cpu().float())
    assert (y - want).abs().max() <= 2 ** -8 * want.abs().max()                             # bf16 output rounding
    dx = ops.gelu_bwd(dy, x).cpu().float()
    xr = x.cpu().float().requires_grad_(True)
    (gx,) = torch.autograd.grad(F.gelu(xr), xr, dy.cpu().float())
    assert (dx - gx).abs().max() <= 2 ** -7 * gx.abs().max()


@pytest.mark.parametrize('rows,gather', [(594, None), (5000, None), (2 * 16, (16, 30, 1)), (3 * 84, (84, 198, 114))])
def test_layernorm_backward(cuda_device, rows, gather):
    torch.manual_seed(rows)
    group, stride, offset = gather if gather else (rows, rows, 0)
    n_dy = (rows // group - 1) * stride + offset + group
    x = torch.randn((rows, D)) * 1.7 + 0.3
    dy = torch.randn((n_dy, D))
    gamma = 1.0 + 0.1 * torch.randn(D)
    dx0 = torch.randn((rows, D))
    want_dx, want_dg, want_db = fake_ops.layernorm_bwd(dy, x, gamma, 1e-5, rows=rows, group=group, group_stride=stride, offset=offset)
    xr = x.clone().requires_grad_(True)
    gr = gamma.clone().requires_grad_(True)
    r = torch.arange(rows)
    ag = torch.autograd.grad(F.layer_norm(xr, (D,), gr, torch.zeros(D), 1e-5), [xr, gr], dy[(r // group) * stride + offset + r % group])
    assert (want_dx - ag[0]).abs().max() < 1e-4 and (want_dg - ag[1]).abs().max() < 1e-3   # the stand-in itself is right
    dev = cuda_device
    got_dx, got_dg, got_db = ops.layernorm_bwd(dy.to(dev), x.to(dev), gamma.to(dev), 1e-5, rows=rows, group=group, group_stride=stride, offset=offset)
    assert (got_dx.cpu() - want_dx).abs().max() < 2e-4
    assert (got_dg.cpu() - want_dg).abs().max() < 1e-3 * max(1.0, float(want_dg.abs().max()))
    assert (got_db.cpu() - want_db).abs().max() < 1e-3 * max(1.0, float(want_db.abs().max()))
    acc = dx0.to(dev).clone()
    ops.layernorm_bwd(dy.to(dev), x.to(dev), gamma.to(dev), 1e-5, dx=acc, accumulate=True, rows=rows, group=group, group_stride=stride, offset=offset)
    assert (acc.cpu() - (dx0 + want_dx)).abs().max() < 2e-4


# ---- attention -----------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('B,T,h,d,p', [(2, 30, 8, 96, 0.0), (2, 30, 8, 96, 0.1), (3, 198, 8, 96, 0.1), (2, 74, 12, 64, 0.1), (1, 257, 8, 96, 0.0)])
def test_attention_train_forward_and_backward(cuda_device, B, T, h, d, p):
    torch.manual_seed(T + d)
    seed, site = 777, 4
    Dm = h * d
    qkv = _bf(torch.randn((B * T, 3 * Dm)) * 0.8)
    d_out = _bf(torch.randn((B * T, Dm)))
    scale = 1.0 / math.sqrt(d)
    want_o, want_lse = fake_ops.attention_train_fwd(qkv.float(), B, T, h, d, scale, p, seed, site)
    dev = cuda_device
    got_o, got_lse = ops.attention_train_fwd(qkv.to(dev), B, T, h, d, scale, p, seed, site)
    assert (got_lse.cpu() - want_lse).abs().max() < 1e-3
    assert (got_o.cpu().float() - want_o).abs().max() <= 2 ** -7 * float(want_o.abs().max()) + 1e-3
    # backward from the GPU's own (bf16) forward output, against the explicit formulas and against autograd of the dense definition
    want_dqkv = fake_ops.attention_train_bwd(qkv.float(), got_o.cpu().float(), d_out.float(), got_lse.cpu(), B, T, h, d, scale, p, seed, site)
    got_dqkv = ops.attention_train_bwd(qkv.to(dev), got_o, d_out.to(dev), got_lse, B, T, h, d, scale, p, seed, site).cpu().float()
    assert (got_dqkv - want_dqkv).abs().max() <= 2 ** -6 * float(want_dqkv.abs().max())
    x = qkv.float().requires_grad_(True)
    q, k, v = x.reshape(B, T, 3, h, d).permute(2, 0, 3, 1, 4)
    m = torch.from_numpy(philox.dropout_multiplier((B, h, T, T), p, seed, site))
    o = ((torch.softmax(q @ k.transpose(-1, -2) * scale, -1) * m) @ v).permute(0, 2, 1, 3).reshape(B * T, Dm)
    (ag,) = torch.autograd.grad(o, x, d_out.float())
    rel = float((got_dqkv - ag).norm() / ag.norm())
    assert rel < 1e-2, rel


def test_attention_train_rejects_unsupported_shapes(cuda_device):
    qkv = _bf(torch.zeros((600, 3 * 768), device=cuda_device))
    with pytest.raises(Exception, match='shared memory|head_dim'):
        ops.attention_train_fwd(qkv, 1, 600, 8, 96, 0.1, 0.0, 0, 0)


# ---- head, linear --------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('B,T,n_cls', [(2, 30, 21), (5, 198, 21), (3, 184, 2)])
def test_sync_head_backward(cuda_device, B, T, n_cls):
    torch.manual_seed(B * T)
    x = torch.randn((B * T, D)) + 0.2
    lw, lb = 1.0 + 0.1 * torch.randn(D), 0.05 * torch.randn(D)
    W, dl = 0.03 * torch.randn((n_cls, D)), torch.randn((B, n_cls))
    want = fake_ops.sync_head_bwd(x, T, lw, lb, 1e-5, W, dl, B)
    leaves = [t.clone().requires_grad_(True) for t in (x, lw, lb, W)]
    bias = torch.zeros(n_cls, requires_grad=True)
    logits = F.linear(F.layer_norm(leaves[0].reshape(B, T, D)[:, 0], (D,), leaves[1], leaves[2], 1e-5), leaves[3], bias)
    ag = torch.autograd.grad(logits, leaves + [bias], dl)
    for w, a in zip(want, ag):
        assert (w - a).abs().max() < 1e-4 * max(1.0, float(a.abs().max()))
    dev = cuda_device
    got = ops.sync_head_bwd(x.to(dev), T, lw.to(dev), lb.to(dev), 1e-5, W.to(dev), dl.to(dev), B)
    for g, a in zip(got, ag):
        assert g.shape == a.shape
        assert (g.cpu() - a).abs().max() < 2e-4 * max(1.0, float(a.abs().max()))


def test_linear_function_gradients(cuda_device):
    torch.manual_seed(3)
    dev = cuda_device
    x = torch.randn((594, D), device=dev, requires_grad=True)          # M % 8 != 0: exercises the zero-padded transposes
    lin = torch.nn.Linear(D, D).to(dev)
    w16 = ops.cast_bf16(lin.weight.detach().contiguous())
    y = train.linear(x, lin.weight, lin.bias, w16)
    dy = torch.randn_like(y)
    gx, gw, gb = torch.autograd.grad(y, [x, lin.weight, lin.bias], dy)
    xr = x.detach().cpu().requires_grad_(True)
    wr, br = lin.weight.detach().cpu().requires_grad_(True), lin.bias.detach().cpu().requires_grad_(True)
    ag = torch.autograd.grad(F.linear(xr, wr, br), [xr, wr, br], dy.cpu())
    assert (y.detach().cpu() - F.linear(xr, wr, br).detach()).abs().max() < 5e-2
    for g, a in zip((gx, gw, gb), ag):
        assert float((g.cpu() - a).norm() / a.norm()) < 1e-2


# ---- the whole step ------------------------------------------------------------------------------------------------------------
def _train_model(S, p_drop, sd, dev, head_target='model.sync_model.GlobalTransformer', n_classes=21):
    cfg = M.sync_yaml_model_config(S, n_classes, head_target)
    for k in ('embd_pdrop', 'resid_pdrop', 'attn_pdrop'):
        cfg['transformer']['params'][k] = p_drop
    cfg = {k: {kk: vv for kk, vv in v.items() if kk != 'is_trainable'} for k, v in cfg.items()}
    model = M.Synchformer(**cfg)
    model.load_state_dict(sd, strict=True)
    model.to(dev).train()
    for ext in (model.vfeat_extractor, model.afeat_extractor):          # scripts/train_utils.py:199-204, 330-342
        ext.requires_grad_(False)
        ext.eval()
    return model


def _step(model, vf, af, targets, seed, monkeypatch, loss_scale=1.0):
    monkeypatch.setattr(train, 'draw_seed', lambda: seed)
    model.zero_grad(set_to_none=True)
    v, a = model.project(vf, af)
    logits = model.transformer(v, a)
    loss = model.compute_loss(logits, targets)
    (loss * loss_scale).backward()
    return loss.detach(), logits.detach(), {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}


@pytest.mark.parametrize('p_drop', [0.0, 0.1])
def test_training_step_gradients_match_oracle_on_golden_features(cuda_device, monkeypatch, p_drop):
    """B = 2, S = 2 (T = 30) on the committed golden features of the frozen extractors; fp32 oracle autograd with the same multipliers.
    Gate: per tensor |grad - ref| <= 3e-2 |ref| + floor (tests/train_gates.py; CPU emulation of the rounding points measures 1.3e-2)."""
    f = np.load(os.path.join(GOLD, 'sync_b2s2.npz'))
    g = np.load(os.path.join(GOLD, 'sync_train_b2s2.npz'))
    B, S, seed_w, seed_drop, _ = (int(x) for x in g['meta'])
    sd = synth.synthetic_state_dict(seed_w, n_segments=S)
    vf, af, targets = torch.from_numpy(f['vfeats']), torch.from_numpy(f['afeats']), torch.from_numpy(f['targets'])
    model = _train_model(S, p_drop, sd, cuda_device)
    loss, logits, grads = _step(model, vf.to(cuda_device), af.to(cuda_device), targets.to(cuda_device), seed_drop, monkeypatch)
    mult = None if p_drop == 0 else O.train_multipliers(B, 2 + 14 * S, seed_drop, p_drop, p_drop, p_drop)
    rloss, rlogits, rgrads = O.sync_train_grads(sd, vf, af, targets, mult)
    assert (logits.cpu() - rlogits).abs().max() < 2e-2
    assert abs(float(loss) - float(rloss)) < 2e-2
    assert len(grads) == 63 and all(n.split('.')[0] in ('vproj', 'aproj', 'transformer') for n in grads)
    worst = train_gates.check_grads(grads, rgrads)
    print(f'p_drop={p_drop}: worst gradient error / allowance {worst:.3f}')
    if p_drop > 0:                                                       # and against the reference's own backward (golden samples)
        stride = int(g['meta'][4])
        for n, gr in grads.items():
            sample = g['drop_sample/' + n]
            assert np.abs(gr.cpu().reshape(-1)[::stride].numpy() - sample).max() <= 5e-2 * np.abs(sample).max() + 2e-3, n


def test_training_step_at_five_second_clip_shape(cuda_device, monkeypatch, check_determinism=True):
    """S = 14 (T = 198), B = 3: B * T = 594 is not a multiple of 8 (padded transposes), 7 row tiles per attention problem."""
    torch.manual_seed(5)
    B, S = 3, 14
    sd = synth.synthetic_state_dict(1337, n_segments=S)
    vf, af = torch.randn((B, S, 8, D)) * 0.5, torch.randn((B, S, 6, D)) * 0.5
    targets = torch.tensor([1, 20, 7])
    model = _train_model(S, 0.1, sd, cuda_device)
    loss, logits, grads = _step(model, vf.to(cuda_device), af.to(cuda_device), targets.to(cuda_device), 4242, monkeypatch, loss_scale=65536.0)
    rloss, rlogits, rgrads = O.sync_train_grads(sd, vf, af, targets, O.train_multipliers(B, 2 + 14 * S, 4242), loss_scale=65536.0)
    assert (logits.cpu() - rlogits).abs().max() < 3e-2
    train_gates.check_grads(grads, rgrads)
    if not check_determinism:
        return
    # same seed -> same step, bit for bit (no atomics anywhere); another seed -> another mask
    _, logits2, grads2 = _step(model, vf.to(cuda_device), af.to(cuda_device), targets.to(cuda_device), 4242, monkeypatch, loss_scale=65536.0)
    assert torch.equal(logits, logits2) and all(torch.equal(grads[n], grads2[n]) for n in grads)
    _, logits3, _ = _step(model, vf.to(cuda_device), af.to(cuda_device), targets.to(cuda_device), 4243, monkeypatch)
    assert not torch.equal(logits, logits3)


def test_syncability_head_trains(cuda_device, monkeypatch):
    """GlobalTransformerWithSyncabilityHead (sync_model.py:176-190): S = 13, 2 classes."""
    torch.manual_seed(6)
    B, S = 2, 13
    sd = synth.synthetic_state_dict(1337, n_segments=S, n_classes=2, head='sync_head')
    model = _train_model(S, 0.0, sd, cuda_device, 'model.sync_model.GlobalTransformerWithSyncabilityHead', 2)
    vf, af, targets = torch.randn((B, S, 8, D)) * 0.5, torch.randn((B, S, 6, D)) * 0.5, torch.tensor([0, 1])
    _, logits, grads = _step(model, vf.to(cuda_device), af.to(cuda_device), targets.to(cuda_device), 1, monkeypatch)
    _, rlogits, rgrads = O.sync_train_grads(sd, vf, af, targets, None, head='sync_head')
    assert logits.shape == (B, 2) and (logits.cpu() - rlogits).abs().max() < 2e-2
    train_gates.check_grads(grads, rgrads)


def test_full_forward_in_train_mode_with_frozen_extractors_and_optimizer_steps(cuda_device):
    """The harness' loop (train_sync.py:177-183, train_utils.py:373-386): model(vid, aud, targets) in train mode under GradScaler,
    clip_grad_norm_, Adam over ALL parameters; the loss on a fixed batch goes down and only the sync module moves."""
    torch.manual_seed(7)
    S = 2
    sd = synth.synthetic_state_dict(1337, n_segments=S)
    model = _train_model(S, 0.1, sd, cuda_device)
    vis = synth.synthetic_video(2, S, 0).to(cuda_device).half()
    aud = O.mel_frontend(synth.synthetic_waveform(2, S, 0)).float().unsqueeze(2).to(cuda_device)
    targets = torch.tensor([3, 17], device=cuda_device)
    opt = torch.optim.Adam(model.parameters(), 2e-4, (0.9, 0.999), 1e-7, 0.0)
    scaler = torch.amp.GradScaler(cuda_device.type, enabled=cuda_device.type == 'cuda')
    params = dict(model.named_parameters())
    frozen, moved = params['vfeat_extractor.blocks.3.attn.qkv.weight'], params['transformer.blocks.1.attn.query.weight']
    frozen_before, moved_before = frozen.detach().clone(), moved.detach().clone()
    losses = []
    for _ in range(6):
        opt.zero_grad(set_to_none=True)
        loss, logits = model(vis, aud, targets)
        scaler.scale(loss).backward()
        scaler.unscale_(opt)
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        scaler.step(opt)
        scaler.update()
        assert torch.isfinite(loss)
        losses.append(float(loss))
    assert all(p.grad is None for p in model.vfeat_extractor.parameters()) and all(p.grad is None for p in model.afeat_extractor.parameters())
    assert torch.equal(frozen_before, frozen.detach())
    assert not torch.equal(moved_before, moved.detach())
    assert min(losses[3:]) < losses[0], losses
    model.eval()                                                        # and the eval path still works on the updated weights
    with torch.no_grad():
        _, lg = model(vis, aud)
    assert torch.isfinite(lg).all()


def test_fused_cross_entropy_and_adam(cuda_device):
    """sfb_cross_entropy vs F.cross_entropy (+ autograd); FusedAdam vs torch.optim.Adam over several steps, plain and with the fused
    unscale + clip_grad_norm_ + skip-on-inf path against GradScaler-style reference arithmetic."""
    from synchformer_b200 import optim
    dev = cuda_device
    torch.manual_seed(0)
    logits = (torch.randn(37, 21) * 3).to(dev).requires_grad_(True)
    targets = torch.randint(0, 21, (37,)).to(dev)
    with torch.enable_grad():
        loss = optim.cross_entropy(logits, targets)
        (g,) = torch.autograd.grad(loss * 65536.0, logits)
        ref = torch.nn.functional.cross_entropy(logits, targets)
        (rg,) = torch.autograd.grad(ref * 65536.0, logits)
    assert abs(float(loss) - float(ref)) < 1e-6 and (g - rg).abs().max() < 1e-6 * 65536

    shapes = [(768, 768), (768,), (3, 70000), (1, 1, 768), (21, 768)]
    def make():
        gen = torch.Generator().manual_seed(1)
        return [torch.nn.Parameter(torch.randn(s, generator=gen).to(dev)) for s in shapes]
    mine, theirs = make(), make()
    a = optim.FusedAdam(mine, lr=1e-2, betas=(0.9, 0.999), eps=1e-7, weight_decay=0.01)
    b = torch.optim.Adam(theirs, lr=1e-2, betas=(0.9, 0.999), eps=1e-7, weight_decay=0.01)
    gen = torch.Generator().manual_seed(2)
    for it in range(4):
        grads = [torch.randn(s, generator=gen) * 0.1 for s in shapes]
        for p, q, gr in zip(mine, theirs, grads):
            p.grad, q.grad = gr.clone().to(dev), gr.clone().to(dev)
        if it < 2:
            a.step()
            b.step()
        else:                      # fused: gradients still carry the loss scale; clip to max_norm 1
            scale = 1024.0
            for p in mine:
                p.grad.mul_(scale)
            norm, found = a.step(grad_scale=scale, max_norm=1.0)
            tn = torch.nn.utils.clip_grad_norm_(theirs, 1.0)
            b.step()
            assert float(found) == 0.0 and abs(float(norm) - float(tn)) < 1e-4 * float(tn)
        for p, q in zip(mine, theirs):
            assert (p - q).abs().max() < 2e-6, it
    # a non-finite gradient skips the step and does not advance the step counter
    before = [p.detach().clone() for p in mine]
    for p in mine:
        p.grad = torch.randn(p.shape, generator=gen).to(dev)
    mine[2].grad[0, 5] = float('inf')
    norm, found = a.step(grad_scale=1.0, max_norm=1.0)
    assert float(found) == 1.0 and all(torch.equal(p, q) for p, q in zip(mine, before))
    assert float(a._dev_state[0]['step']) == 4.0
    # every real step bumps the parameters' version counters (the bf16 weight caches of model.py are keyed on them)
    assert all(p._version >= 4 for p in mine)
    # the step counter round-trips through state_dict(), in both directions between FusedAdam and torch.optim.Adam (resume from a checkpoint
    # that either optimiser wrote): one more step after the reload must match the uninterrupted torch.optim.Adam run
    assert all(float(a.state[p]['step']) == 4.0 for p in mine)
    grads = [torch.randn(s, generator=gen) * 0.1 for s in shapes]
    resumed = {}
    for name, src in (('fused->fused', a), ('adam->fused', b)):
        ps = [torch.nn.Parameter(q.detach().clone()) for q in theirs]
        opt = optim.FusedAdam(ps, lr=1e-2, betas=(0.9, 0.999), eps=1e-7, weight_decay=0.01)
        opt.load_state_dict(copy.deepcopy(src.state_dict()))      # load_state_dict aliases tensors that already have the right dtype / device
        for p, gr in zip(ps, grads):
            p.grad = gr.clone().to(dev)
        opt.step()
        assert float(opt.state[ps[0]]['step']) == 5.0, name
        resumed[name] = ps
    ps = [torch.nn.Parameter(q.detach().clone()) for q in theirs]
    opt = torch.optim.Adam(ps, lr=1e-2, betas=(0.9, 0.999), eps=1e-7, weight_decay=0.01)
    opt.load_state_dict(copy.deepcopy(a.state_dict()))               # fused -> torch.optim.Adam
    for p, q, gr in zip(ps, theirs, grads):
        p.grad, q.grad = gr.clone().to(dev), gr.clone().to(dev)
    opt.step()
    b.step()
    for name, got in list(resumed.items()) + [('fused->adam', ps)]:
        for p, q in zip(got, theirs):
            assert (p - q).abs().max() < 2e-6, name
    # F.cross_entropy target conventions: -100 rows are ignored (loss, gradient, denominator); other out-of-range targets poison the loss
    t2 = targets.clone()
    t2[::5] = -100
    lg2 = logits.detach().clone().requires_grad_(True)
    with torch.enable_grad():
        loss = optim.cross_entropy(lg2, t2)
        (g,) = torch.autograd.grad(loss, lg2)
        ref = torch.nn.functional.cross_entropy(lg2, t2)
        (rg,) = torch.autograd.grad(ref, lg2)
    assert abs(float(loss) - float(ref)) < 1e-6 and (g - rg).abs().max() < 1e-6 and float(g[0].abs().max()) == 0.0
    t2[1] = 21
    assert torch.isnan(optim.cross_entropy(logits.detach(), t2))


def test_fused_adam_updates_reach_the_model_forward(cuda_device, monkeypatch):
    """FusedAdam writes parameters through raw pointers; the bf16 GEMM-weight caches are keyed on the parameters' version
    counters, so every step must bump them.  Three training steps of the synchronisation module with FusedAdam must track the same steps
    with torch.optim.Adam: same losses / logits to bf16 noise, and the logits must MOVE between steps (a stale cache would freeze the GEMM weights)."""
    from synchformer_b200 import optim
    B, S = 2, 2
    torch.manual_seed(11)
    sd = synth.synthetic_state_dict(1337, n_segments=S)
    vf, af = (torch.randn((B, S, 8, D)) * 0.5).to(cuda_device), (torch.randn((B, S, 6, D)) * 0.5).to(cuda_device)
    targets = torch.tensor([3, 17]).to(cuda_device)
    runs = {}
    for kind in ('fused', 'torch'):
        model = _train_model(S, 0.0, sd, cuda_device)
        params = [p for n, p in model.named_parameters() if n.startswith(('vproj', 'aproj', 'transformer'))]
        opt = (optim.FusedAdam if kind == 'fused' else torch.optim.Adam)(params, lr=3e-3, betas=(0.9, 0.999), eps=1e-7)
        trace = []
        for it in range(3):
            opt.zero_grad(set_to_none=True)
            v, a = model.project(vf, af)
            logits = model.transformer(v, a)
            loss = model.compute_loss(logits, targets)
            loss.backward()
            opt.step()
            trace.append((float(loss), logits.detach().float().cpu()))
        model.eval()
        with torch.no_grad():
            v, a = model.project(vf, af)
            trace.append((0.0, model.transformer(v, a).float().cpu()))
        runs[kind] = trace
    f, t = runs['fused'], runs['torch']
    assert (f[0][1] - f[1][1]).abs().max() > 1e-2 and (f[1][1] - f[2][1]).abs().max() > 1e-3, 'logits did not move: stale bf16 weight cache'
    for (lf, gf), (lt, gt) in zip(f, t):
        assert abs(lf - lt) < 2e-2 * max(1.0, abs(lt)), (lf, lt)
        assert (gf - gt).abs().max() < 5e-2 * max(1.0, float(gt.abs().max())), float((gf - gt).abs().max())


def test_gradients_scale_exactly_with_the_loss_scale_at_config4_size(cuda_device, monkeypatch):
    """Size-independent property at BASELINE config 4's per-GPU size (32 clips x 14 segments, T = 198, M = 6336 rows): the backward is
    linear in the upstream gradient, and a power-of-two loss scale (what GradScaler applies) commutes with every bf16 / fp32 rounding in
    it, so gradients under scale 2^12 are bit-for-bit 2^12 times the unscaled ones.  Same seed -> same masks."""
    B, S = (32, 14) if cuda_device.type == 'cuda' else (2, 3)
    torch.manual_seed(8)
    sd = synth.synthetic_state_dict(1337, n_segments=S)
    model = _train_model(S, 0.1, sd, cuda_device)
    vf, af = (torch.randn((B, S, 8, D)) * 0.5).to(cuda_device), (torch.randn((B, S, 6, D)) * 0.5).to(cuda_device)
    targets = torch.randint(0, 21, (B,)).to(cuda_device)
    _, logits1, g1 = _step(model, vf, af, targets, 99, monkeypatch, loss_scale=1.0)
    _, logits2, g2 = _step(model, vf, af, targets, 99, monkeypatch, loss_scale=4096.0)
    assert torch.equal(logits1, logits2)
    assert all(torch.isfinite(g).all() for g in g2.values())
    for n in g1:
        assert torch.equal(g1[n] * 4096.0, g2[n]), n
