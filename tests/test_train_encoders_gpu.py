"""GPU parity of the encoders' training step: the new kernels against autograd of their dense definitions, both towers
(forward with DropPath + hand-written backward, 448 gradient tensors) against torch autograd on the fp32 CPU oracle with the same DropPath
multipliers, the AVCLIP step and stage II with trainable extractors end to end.

The same kernel sources also pass these checks on the CPU SIMT emulator (tests/test_train_encoders_cpu.py, tests/emu/).  These tests gate:
they were seen green on a B200 (GPUTEST_r01) and carry no xfail marker.
"""
import os

import numpy as np
import pytest
import torch

from oracle import philox
from oracle import synchformer_oracle as O
from synchformer_b200 import avclip, model as M, ops, synth, train_encoders as TE

import test_train_encoders_cpu as C

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _grad_enabled():
    with torch.enable_grad():
        yield


@pytest.fixture(scope='module')
def setup():
    g = np.load(os.path.join(C.HERE, 'golden', 'encoders_train_b1s2.npz'))
    B, S, seed_w, seed_x, seed_drop, stride = (int(x) for x in g['meta'])
    sd = synth.synthetic_state_dict(seed_w, n_segments=S)
    vis = synth.synthetic_video(B, S, seed_x)
    aud = O.mel_frontend(synth.synthetic_waveform(B, S, seed_x)).float().unsqueeze(2)
    ref = O.encoders_train_grads(sd, vis, aud, O.drop_path_multipliers(B * S, seed_drop))
    return dict(g=g, B=B, S=S, sd=sd, vis=vis, aud=aud, seed=seed_drop, stride=stride, ref=ref)


@pytest.mark.parametrize('mode', ['space', 'time'])
def test_divided_attention_backward(cuda_device, mode):
    torch.manual_seed(3)
    n, D, TOK, h, d = 2, 768, 1569, 12, 64
    qkv = (torch.randn(n * TOK, 3 * D) * 0.7).to(torch.bfloat16)
    d_o = torch.randn(n * TOK, D).to(torch.bfloat16)
    x = qkv.float().requires_grad_(True)
    q, k, v = [t.reshape(n, TOK, h, d).permute(0, 2, 1, 3) for t in x.chunk(3, -1)]
    cls = torch.softmax(q[:, :, 0:1] @ k.transpose(-1, -2) * 0.125, -1) @ v
    re = (lambda t: t.reshape(n, h, 8, 196, d).permute(0, 1, 3, 2, 4)) if mode == 'time' else (lambda t: t.reshape(n, h, 8, 196, d))
    q_, k_, v_ = re(q[:, :, 1:]), re(k[:, :, 1:]), re(v[:, :, 1:])
    G = q_.shape[2]
    ck, cv = k[:, :, 0:1].unsqueeze(2).expand(n, h, G, 1, d), v[:, :, 0:1].unsqueeze(2).expand(n, h, G, 1, d)
    out = torch.softmax(q_ @ torch.cat([ck, k_], 3).transpose(-1, -2) * 0.125, -1) @ torch.cat([cv, v_], 3)
    out = (out.permute(0, 1, 3, 2, 4) if mode == 'time' else out).reshape(n, h, 1568, d)
    out = torch.cat([cls, out], 2).permute(0, 2, 1, 3).reshape(n * TOK, D)
    (ref,) = torch.autograd.grad(out, x, d_o.float())
    dev = cuda_device
    got = TE._divided_attention_bwd(qkv.to(dev), out.detach().to(torch.bfloat16).to(dev), d_o.to(dev), n, mode).float().cpu()
    assert float((got - ref).norm() / ref.norm()) < 5e-3
    rows0 = torch.arange(n) * TOK
    assert float((got[rows0] - ref[rows0]).norm() / ref[rows0].norm()) < 5e-3
    again = TE._divided_attention_bwd(qkv.to(dev), out.detach().to(torch.bfloat16).to(dev), d_o.to(dev), n, mode).float().cpu()
    assert torch.equal(got, again)                                   # no atomics: bit-reproducible


def test_droppath_and_row_gather(cuda_device):
    dev = cuda_device
    x, res = torch.randn(6 * 1569, 768), torch.randn(6 * 1569, 768)
    m = torch.from_numpy(philox.dropout_multiplier((6,), 0.4, 11, 9)).repeat_interleave(1569).unsqueeze(1)
    assert torch.equal(ops.droppath(x.to(dev), 1569, 0.4, 11, 9, residual=res.to(dev)).cpu(), res + x * m)
    assert torch.equal(ops.droppath(x.to(dev), 1569, 0.4, 11, 9, out_bf16=True).cpu(), (x * m).to(torch.bfloat16))
    r = torch.arange(6 * 1568)
    assert torch.equal(ops.gather_rows_bf16(x.to(dev), 6 * 1568, 1568, 1569, 1).cpu(), x[(r // 1568) * 1569 + 1 + r % 1568].to(torch.bfloat16))


def _gpu_grads(s, dev, towers):
    model = M.build_synchformer(n_segments=s['S'], state_dict=s['sd'], device=dev)
    model.train()
    _, rv, ra, _ = s['ref']
    vf = TE.motionformer_features(model.vfeat_extractor, s['vis'].to(dev), seed=s['seed']) if 'v' in towers else rv.clone().to(dev)
    af = TE.ast_features(model.afeat_extractor, s['aud'].view(s['B'], s['S'], 128, 66).to(dev)) if 'a' in towers else ra.clone().to(dev)
    v = torch.nn.functional.normalize(vf.mean(2).reshape(-1, 768), dim=-1)
    a = torch.nn.functional.normalize(af.mean(2).reshape(-1, 768), dim=-1)
    tgt = torch.eye(v.shape[0], device=dev)
    loss = (torch.nn.functional.cross_entropy(v @ a.mT / 0.07, tgt) + torch.nn.functional.cross_entropy(a @ v.mT / 0.07, tgt)) / 2
    loss.backward()
    grads = {}
    for pref, mod in (('vfeat_extractor.', model.vfeat_extractor), ('afeat_extractor.', model.afeat_extractor)):
        grads.update({pref + n: p.grad.cpu() for n, p in mod.named_parameters() if p.grad is not None})
    return loss.detach().cpu(), vf.detach().cpu(), af.detach().cpu(), grads


def test_ast_tower_gradients_match_oracle(cuda_device, setup):
    loss, vf, af, grads = _gpu_grads(setup, cuda_device, ('a',))
    _, rv, ra, rg = setup['ref']
    assert float((af - ra).norm() / ra.norm()) < 1e-2
    C._check_bf16(grads, rg, 'afeat_extractor.')


def test_motionformer_tower_gradients_match_oracle(cuda_device, setup):
    loss, vf, af, grads = _gpu_grads(setup, cuda_device, ('v',))
    _, rv, ra, rg = setup['ref']
    assert float((vf - rv).norm() / rv.norm()) < 1e-2
    C._check_bf16(grads, rg, 'vfeat_extractor.')
    # against the reference's own backward as well (golden samples)
    g, stride = setup['g'], setup['stride']
    for n, gr in grads.items():
        if n.startswith('vfeat_extractor.') and 'key' not in n:
            sample = g['sample/' + n]
            assert np.abs(gr.reshape(-1)[::stride].numpy() - sample).max() <= 0.25 * np.abs(sample).max() + 1e-5, n


def test_avclip_training_steps(cuda_device):
    """stage I: AVCLIP(...).train(); loss.backward(); AdamW step - the loss on a fixed batch goes down, every tower parameter that takes part
    in the forward gets a gradient, eval mode still runs the inference kernels."""
    torch.manual_seed(0)
    dev = cuda_device
    model = avclip.AVCLIP().to(dev).train()
    sd = synth.synthetic_state_dict(1337, n_segments=2)
    model.v_encoder.load_state_dict({k[len('vfeat_extractor.'):]: v for k, v in sd.items() if k.startswith('vfeat_extractor.')})
    model.a_encoder.load_state_dict({k[len('afeat_extractor.'):]: v for k, v in sd.items() if k.startswith('afeat_extractor.')})
    vis = synth.synthetic_video(2, 2, 0).permute(0, 1, 3, 2, 4, 5).contiguous().to(dev)                 # stage-I layout (B, S, C, T, H, W)
    aud = O.mel_frontend(synth.synthetic_waveform(2, 2, 0)).float().permute(0, 1, 3, 2).contiguous().to(dev)   # (B, S, T, F)
    opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-4)
    losses = []
    for _ in range(4):
        opt.zero_grad(set_to_none=True)
        loss = model(vis, aud)['losses']['segment_contrastive_loss']
        loss.backward()
        opt.step()
        assert torch.isfinite(loss)
        losses.append(float(loss))
    missing = [n for n, p in model.named_parameters() if p.requires_grad and p.grad is None]
    assert missing == [], missing
    assert min(losses[1:]) < losses[0], losses
    model.eval()
    with torch.no_grad():
        out = model(vis, aud)
    assert torch.isfinite(out['losses']['segment_contrastive_loss'])


def test_stage_two_with_trainable_extractors(cuda_device):
    """configs/sync.yaml with is_trainable: True: gradients reach the towers through vproj / aproj and the sync transformer."""
    S = 2
    dev = cuda_device
    model = M.build_synchformer(n_segments=S, state_dict=synth.synthetic_state_dict(1337, n_segments=S), device=dev)
    model.train()
    vis = synth.synthetic_video(1, S, 0).to(dev).half()
    aud = O.mel_frontend(synth.synthetic_waveform(1, S, 0)).float().unsqueeze(2).to(dev)
    loss, logits = model(vis, aud, torch.tensor([4], device=dev))
    loss.backward()
    got = {n for n, p in model.named_parameters() if p.grad is not None}
    want = {n for n, p in model.named_parameters() if '.patch_embed.proj.' not in n}
    assert got == want, (want - got, got - want)
    assert all(torch.isfinite(p.grad).all() for p in model.parameters() if p.grad is not None)


def test_shifted_window_eval(cuda_device):
    """N4: avclip.shift_and_get_preds on the GPU GEMM against the reference formulation; ties are avoided by construction, and the
    predictions must agree wherever the fp32 similarity margin exceeds the bf16 operand noise."""
    torch.manual_seed(0)
    B, S, W = 16, 14, 8
    a = torch.nn.functional.normalize(torch.randn(B, S, 768), dim=-1)
    v = torch.nn.functional.normalize(a + 0.3 * torch.randn(B, S, 768), dim=-1)          # in-sync tracks: the diagonal wins clearly
    pa, pv = avclip.shift_and_get_preds(a.to(cuda_device), v.to(cuda_device), W)
    ra, rv, sim = O.shift_and_get_preds(a, v, W)
    top2 = sim.topk(2, dim=-1).values
    clear = (top2[..., 0] - top2[..., 1]) > 0.05
    assert clear.float().mean() > 0.9
    assert torch.equal(pv.cpu()[clear], rv[clear])
    gt = torch.arange(S - W + 1).view(1, -1)
    assert (pv.cpu() == gt).float().mean() > 0.9 and (pa.cpu() == gt).float().mean() > 0.9


@pytest.mark.parametrize('n,N', [(6, 6), (5, 15), (128, 128)])
def test_contrastive_tail_kernels_match_oracle_autograd(cuda_device, n, N):
    """csrc/contrastive.cu (mean pooling, L2 normalise, similarities + symmetric CE, backward) against torch autograd on the oracle's
    restatement of open_clip/model.py:507-545; N > n stands for gathered features of other ranks (eye(n, N) targets as in the reference)."""
    dev = cuda_device
    if n > 16 and dev.type != 'cuda':
        pytest.skip('large case on hardware only')
    g = torch.Generator().manual_seed(n * 31 + N)
    T, Dm = 8, 768
    xv = torch.randn(n, T, Dm, generator=g, requires_grad=True)
    xa = torch.randn(n, 6, Dm, generator=g, requires_grad=True)
    others_v = torch.nn.functional.normalize(torch.randn(N - n, Dm, generator=g), dim=-1)
    others_a = torch.nn.functional.normalize(torch.randn(N - n, Dm, generator=g), dim=-1)
    scale = torch.tensor(0.07, requires_grad=True)
    # oracle
    vn, an = torch.nn.functional.normalize(xv.mean(1), dim=-1), torch.nn.functional.normalize(xa.mean(1), dim=-1)
    v_all, a_all = torch.cat([vn, others_v]), torch.cat([an, others_a])
    ref = O.avclip_loss(vn, an, scale, v_all if N > n else None, a_all if N > n else None)
    ref_grads = torch.autograd.grad(ref * 3.0, [xv, xa, scale])
    # kernels
    dxv, dxa = xv.detach().to(dev).requires_grad_(True), xa.detach().to(dev).requires_grad_(True)
    dscale = scale.detach().to(dev).requires_grad_(True)
    kvn, kan = avclip._L2Normalize.apply(M._MeanTokens.apply(dxv)), avclip._L2Normalize.apply(M._MeanTokens.apply(dxa))
    assert (kvn.detach().cpu() - vn.detach()).abs().max() < 1e-6
    if N > n:
        kv_all, ka_all = torch.cat([kvn, others_v.to(dev)]), torch.cat([kan, others_a.to(dev)])
        loss = avclip._ContrastiveLoss.apply(kvn, kan, kv_all, ka_all, dscale)
    else:
        loss = avclip._ContrastiveLoss.apply(kvn, kan, None, None, dscale)
    grads = torch.autograd.grad(loss * 3.0, [dxv, dxa, dscale])
    assert abs(float(loss) - float(ref)) < 2e-5 * max(1.0, abs(float(ref)))
    for got, want, name in zip(grads, ref_grads, ('d video tokens', 'd audio tokens', 'd logit_scale')):
        err = float((got.cpu() - want).abs().max())
        assert err <= 2e-4 * float(want.abs().max()) + 1e-7, (name, err, float(want.abs().max()))
