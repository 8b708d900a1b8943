"""CPU: stage-I training step of the two encoders.
  * the oracle (autograd on the CPU restatement, explicit DropPath multipliers) against the golden made from the live reference towers
  * the host orchestration of synchformer_b200/train_encoders.py on fp32 CPU stand-ins == oracle, all 448 gradient tensors
  * the REAL kernel sources on the CPU SIMT emulator (tests/emu): attention backward kernels against autograd, the AST tower end to end;
    the Motionformer tower end to end takes ~5 min of emulation and runs with SFB_EMU_FULL=1
"""
import os
import statistics

import numpy as np
import pytest
import torch

from oracle import synchformer_oracle as O
from synchformer_b200 import model as M, synth

import fake_ops
from emu import binding

HERE = os.path.dirname(os.path.abspath(__file__))
FLOOR = 1e-8          # the key biases have mathematically zero gradients (softmax shift invariance): pure round-off


@pytest.fixture(autouse=True)
def _grad_enabled():
    with torch.enable_grad():
        yield


@pytest.fixture(scope='module')
def setup():
    g = np.load(os.path.join(HERE, 'golden', 'encoders_train_b1s2.npz'))
    B, S, seed_w, seed_x, seed_drop, stride = (int(x) for x in g['meta'])
    sd = synth.synthetic_state_dict(seed_w, n_segments=S)
    vis = synth.synthetic_video(B, S, seed_x)
    aud = O.mel_frontend(synth.synthetic_waveform(B, S, seed_x)).float().unsqueeze(2)
    ref = O.encoders_train_grads(sd, vis, aud, O.drop_path_multipliers(B * S, seed_drop))
    return dict(g=g, B=B, S=S, sd=sd, vis=vis, aud=aud, seed=seed_drop, stride=stride, ref=ref)


def test_oracle_matches_reference_golden(setup):
    """oracle autograd vs the reference towers' own backward (tests/golden/make_golden_encoders_train.py), fp32 both"""
    s = setup
    g = s['g']
    loss, v, a, grads = s['ref']
    assert abs(float(loss) - float(g['loss'])) < 2e-6
    assert np.abs(v.numpy() - g['vfeats']).max() < 5e-5 and np.abs(a.numpy() - g['afeats']).max() < 5e-5
    assert len(grads) == 448
    for n, gr in grads.items():
        flat = gr.double().reshape(-1)
        stat, sample = g['stat/' + n], g['sample/' + n]
        assert abs(float(flat.norm()) - stat[0]) <= 1e-3 * stat[0] + FLOOR, n
        err = np.abs(flat[::s['stride']].float().numpy() - sample).max()
        assert err <= 1e-3 * np.abs(sample).max() + FLOOR, (n, err)


def _product_grads(setup, towers=('v', 'a')):
    from synchformer_b200 import train_encoders as TE
    s = setup
    model = M.build_synchformer(n_segments=s['S'], state_dict=s['sd'])
    model.train()
    _, rv, ra, _ = s['ref']
    vf = TE.motionformer_features(model.vfeat_extractor, s['vis'], seed=s['seed']) if 'v' in towers else rv.clone()
    af = TE.ast_features(model.afeat_extractor, s['aud'].view(s['B'], s['S'], 128, 66)) if 'a' in towers else ra.clone()
    loss = O.contrastive_loss(vf, af, 0.07)
    loss.backward()
    grads = {}
    for pref, mod in (('vfeat_extractor.', model.vfeat_extractor), ('afeat_extractor.', model.afeat_extractor)):
        grads.update({pref + n: p.grad for n, p in mod.named_parameters() if p.grad is not None})
    return loss.detach(), vf.detach(), af.detach(), grads


def test_host_orchestration_matches_oracle_exactly_in_fp32(setup, monkeypatch):
    """train_encoders.py on fp32 stand-ins for every kernel == oracle autograd: all operands / transposes / gathers / DropPath sites /
    shared-CLS reductions of both towers are right (448 tensors, <= 1e-4 relative)."""
    fake_ops.install(monkeypatch, round_bf16=False, names=fake_ops.ALL + fake_ops.ENCODER_FWD + fake_ops.N1_BWD)
    loss, vf, af, grads = _product_grads(setup)
    rloss, rv, ra, rg = setup['ref']
    assert abs(float(loss) - float(rloss)) < 1e-5
    assert (vf - rv).abs().max() < 1e-4 and (af - ra).abs().max() < 1e-4
    assert set(grads) == set(rg)
    for n, r in rg.items():
        err, ref = float((grads[n].double() - r.double()).norm()), float(r.double().norm())
        assert err <= 1e-4 * ref + FLOOR, (n, err, ref)


def _check_bf16(grads, rg, prefix):
    """gate for the bf16 path: per tensor |g - ref| <= 0.15 |ref| + floor, median <= 0.05 (the reference's own bf16-autocast backward
    scores median 0.06 / max 0.09 against its fp32 backward on this step - two segments and a 1 / 0.07 logit scale amplify rounding)"""
    ref = {k: v for k, v in rg.items() if k.startswith(prefix)}
    assert {k for k in grads if k.startswith(prefix)} == set(ref)
    floor = 2e-2 * statistics.median(float(r.double().norm()) for r in ref.values())
    rels = []
    for n, r in ref.items():
        err, rn = float((grads[n].double() - r.double()).norm()), float(r.double().norm())
        assert err <= 0.15 * rn + floor, (n, err, rn)
        rels.append(err / (rn + floor))
    assert statistics.median(rels) < 0.05, statistics.median(rels)


def test_ast_tower_on_emulated_kernels(setup, monkeypatch):
    binding.install(monkeypatch)
    loss, vf, af, grads = _product_grads(setup, towers=('a',))
    rloss, rv, ra, rg = setup['ref']
    assert float((af - ra).norm() / ra.norm()) < 1e-2
    _check_bf16(grads, rg, 'afeat_extractor.')


@pytest.mark.skipif(os.environ.get('SFB_EMU_FULL') != '1', reason='~12 min of SIMT emulation (all kernels but the GEMM from real sources); SFB_EMU_FULL=1 runs it (passed when written)')
def test_motionformer_tower_on_emulated_kernels(setup, monkeypatch):
    binding.install(monkeypatch)
    loss, vf, af, grads = _product_grads(setup, towers=('v',))
    rloss, rv, ra, rg = setup['ref']
    assert float((vf - rv).norm() / rv.norm()) < 1e-2
    _check_bf16(grads, rg, 'vfeat_extractor.')


@pytest.mark.parametrize('mode', ['space', 'time'])
def test_divided_attention_backward_kernels_on_emulator(monkeypatch, mode):
    """sfb_attention_bwd (shared CLS prefix key) + sfb_colsum + sfb_attention_bwd_global_query on the Motionformer layout against autograd of
    the dense definition of DividedAttention.forward (vit_helper.py:100-158)."""
    binding.install(monkeypatch)
    from synchformer_b200 import train_encoders as TE
    torch.manual_seed(3)
    n, D, TOK, h, d = 1, 768, 1569, 12, 64
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
    got = TE._divided_attention_bwd(qkv, out.detach().to(torch.bfloat16), d_o, n, mode).float()
    assert float((got - ref).norm() / ref.norm()) < 5e-3
    assert float((got[0] - ref[0]).norm() / ref[0].norm()) < 5e-3            # the CLS row: dq of the global query, dK / dV summed over everything


def test_droppath_and_gather_on_emulator(monkeypatch):
    binding.install(monkeypatch)
    from oracle import philox
    from synchformer_b200 import ops
    x, res = torch.randn(6 * 5, 768), torch.randn(6 * 5, 768)
    m = torch.from_numpy(philox.dropout_multiplier((6,), 0.4, 11, 9)).repeat_interleave(5).unsqueeze(1)
    assert 0 < int((m == 0).sum()) < 30
    assert torch.equal(ops.droppath(x, 5, 0.4, 11, 9, residual=res), res + x * m)
    assert torch.equal(ops.droppath(x, 5, 0.4, 11, 9, out_bf16=True), (x * m).to(torch.bfloat16))
    assert torch.equal(ops.droppath(x, 5, 0.0, 11, 9), x)
    r = torch.arange(4 * 7)
    assert torch.equal(ops.gather_rows_bf16(x, 4 * 7 // 7 * 4, 4, 5, 1)[:16], x[(torch.arange(16) // 4) * 5 + 1 + torch.arange(16) % 4].to(torch.bfloat16))


def test_avclip_train_mode_routes_through_the_differentiable_towers(monkeypatch):
    from synchformer_b200 import avclip, train_encoders as TE
    calls = []
    monkeypatch.setattr(TE, 'motionformer_features', lambda m, vis: (calls.append('v'), torch.zeros(vis.shape[0], vis.shape[1], 768, requires_grad=True))[1])
    monkeypatch.setattr(TE, 'ast_features', lambda m, spec: (calls.append('a'), torch.ones(spec.shape[0], spec.shape[1], 768, requires_grad=True))[1])
    fake_ops.install(monkeypatch, names=('require_cuda',) + fake_ops.CONTRASTIVE)
    model = avclip.AVCLIP().train()
    out = model(torch.zeros(1, 2, 3, 16, 224, 224), torch.zeros(1, 2, 66, 128))
    assert calls == ['v', 'a'] and out['losses']['segment_contrastive_loss'].requires_grad
    calls.clear()
    model.eval()
    with pytest.raises(Exception):            # eval keeps the inference kernels, which need the GPU library: never the differentiable path
        model(torch.zeros(1, 2, 3, 16, 224, 224), torch.zeros(1, 2, 66, 128))
    assert calls == []


@pytest.mark.parametrize('B,S,W', [(3, 14, 8), (2, 5, 5), (1, 7, 1)])
def test_shifted_window_eval_host_logic(monkeypatch, B, S, W):
    """avclip.shift_and_get_preds (one GEMM + index plumbing) == the reference's unfold + bmm formulation (training/train.py:549-579)."""
    from synchformer_b200 import avclip
    fake_ops.install(monkeypatch, round_bf16=False, names=('require_cuda', 'cast_bf16', 'gemm'))
    torch.manual_seed(S)
    a, v = torch.randn(B, S, 768), torch.randn(B, S, 768)
    v = v + 0.5 * a.roll(1, dims=1)                              # some structure so that the arg-max is not a coin flip
    pa, pv = avclip.shift_and_get_preds(a, v, W)
    ra, rv, _ = O.shift_and_get_preds(a, v, W)
    assert pa.shape == (B, S - W + 1) and torch.equal(pa, ra) and torch.equal(pv, rv)


@pytest.mark.parametrize('Lq,Lk,prefix', [(196, 196, True), (64, 100, True), (130, 77, False)])
def test_mma_attention_backward_matches_cuda_core_kernels_and_autograd(monkeypatch, Lq, Lk, prefix):
    """sfb_attention_bwd: the mma.sync kernel (large head-dim-64 problems) against the CUDA-core kernel pair (impl = 1) and against autograd of
    the dense definition, on the emulator; sizes that are not multiples of 16 exercise the padding / masking."""
    lib = binding.install(monkeypatch)
    from synchformer_b200 import ops
    torch.manual_seed(Lq + Lk)
    D, hd, heads, n_inner = 768, 64, 2, 2
    rows = n_inner * max(Lq, Lk) + 1
    qkv = (torch.randn(rows, 3 * D) * 0.7).to(torch.bfloat16)
    d_o = torch.randn(rows, D).to(torch.bfloat16)
    row = 3 * D
    kw = dict(q_strides=(rows * row, Lq * row, row), kv_strides=(rows * row, Lk * row, row), o_strides=(rows * D, Lq * D, D), n_outer=1, n_inner=n_inner,
              n_heads=heads, head_dim=hd, Lq=Lq, Lk=Lk, scale=0.125)
    pk = dict(k_prefix=qkv[:, D:], v_prefix=qkv[:, 2 * D:], prefix_outer=rows * row) if prefix else {}
    att = torch.zeros(rows, D, dtype=torch.bfloat16)
    monkeypatch.setattr(fake_ops, 'REAL_DTYPES', True)
    fake_ops.attention(qkv[1:], qkv[1:, D:], qkv[1:, 2 * D:], att[1:], **kw, **pk)
    res = {}
    for name, impl in (('mma', 0), ('cuda', 1)):
        dqkv = torch.zeros_like(qkv)
        before = lib.emu_launch_count()
        part = ops.attention_bwd(qkv[1:], qkv[1:, D:], qkv[1:, 2 * D:], att[1:], d_o[1:], dqkv[1:], dqkv[1:, D:], dqkv[1:, 2 * D:], impl=impl, **kw, **pk)
        res[name] = (dqkv, part, lib.emu_launch_count() - before)
    assert res['mma'][2] == 1 and res['cuda'][2] == 2                     # one fused launch vs the two-pass pair
    ref = torch.zeros_like(qkv)
    ref_part = fake_ops.attention_bwd(qkv[1:], qkv[1:, D:], qkv[1:, 2 * D:], att[1:], d_o[1:], ref[1:], ref[1:, D:], ref[1:, 2 * D:], **kw, **pk)
    cols = torch.cat([torch.arange(heads * hd) + k * D for k in range(3)])
    r = ref[:, cols].float()
    for name in ('mma', 'cuda'):
        g = res[name][0][:, cols].float()
        assert float((g - r).norm() / r.norm()) < 6e-3, name
        if prefix:
            assert float((res[name][1] - ref_part).norm() / ref_part.norm()) < 6e-3, name
    assert float((res['mma'][0][:, cols].float() - res['cuda'][0][:, cols].float()).norm() / r.norm()) < 6e-3


@pytest.mark.parametrize('n,N', [(6, 6), (5, 15)])
def test_contrastive_tail_kernels_on_the_emulator(monkeypatch, n, N):
    """the real sources of csrc/contrastive.cu on the CPU SIMT emulator, checked by the same test body that runs on the B200"""
    import test_train_encoders_gpu as G
    binding.install(monkeypatch)
    with torch.enable_grad():
        G.test_contrastive_tail_kernels_match_oracle_autograd(torch.device('cpu'), n, N)


def _gather_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    mp = pytest.MonkeyPatch()
    fake_ops.install(mp, names=('require_cuda',) + fake_ops.CONTRASTIVE)
    from synchformer_b200 import avclip
    g = torch.Generator().manual_seed(5)
    v_all, a_all = torch.randn(world * 3, 768, generator=g), torch.randn(world * 3, 768, generator=g)
    v = v_all[rank * 3:(rank + 1) * 3].clone().requires_grad_(True)
    a = a_all[rank * 3:(rank + 1) * 3].clone().requires_grad_(True)
    scale = torch.tensor(0.07, requires_grad=True)
    with torch.enable_grad():
        vn, an = avclip._L2Normalize.apply(v), avclip._L2Normalize.apply(a)
        loss = avclip._ContrastiveLoss.apply(vn, an, avclip._AllGatherRows.apply(vn, None), avclip._AllGatherRows.apply(an, None), scale)
        loss.backward()
    q.put((rank, loss.detach(), v.grad, a.grad, scale.grad))
    dist.barrier()
    dist.destroy_process_group()
    mp.undo()


def test_gather_for_loss_matches_the_reference_formulation_on_two_ranks():
    """gather_for_loss=True (open_clip/model.py:492-494) under world_size-2 gloo: every rank's loss and gradients equal torch autograd on
    the oracle's restatement where the gathered features are differentiable on EVERY rank (torch.distributed.nn.all_gather semantics: the
    backward of the gather sums the ranks' gradients for each block)."""
    import torch.multiprocessing as tmp
    world = 2
    ctx = tmp.get_context('spawn')
    q = ctx.Queue()
    port = 33500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gather_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = {r[0]: r[1:] for r in (q.get(timeout=240) for _ in procs)}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = torch.Generator().manual_seed(5)
    v_all = torch.randn(world * 3, 768, generator=g).requires_grad_(True)
    a_all = torch.randn(world * 3, 768, generator=g).requires_grad_(True)
    scale = torch.tensor(0.07, requires_grad=True)
    with torch.enable_grad():
        vn, an = torch.nn.functional.normalize(v_all, dim=-1), torch.nn.functional.normalize(a_all, dim=-1)
        losses = [O.avclip_loss(vn[r * 3:(r + 1) * 3], an[r * 3:(r + 1) * 3], scale, vn, an) for r in range(world)]
        sum(losses).backward()               # DDP averages afterwards; the sum over ranks is what the per-rank backward passes add up to
    for r in range(world):
        loss, gv, ga, gs = outs[r]
        assert abs(float(loss) - float(losses[r])) < 1e-5
        assert (gv - v_all.grad[r * 3:(r + 1) * 3]).abs().max() < 1e-5 * max(1.0, float(v_all.grad.abs().max()))
        assert (ga - a_all.grad[r * 3:(r + 1) * 3]).abs().max() < 1e-5 * max(1.0, float(a_all.grad.abs().max()))
    # d logit_scale is local to each rank's loss (DDP all-reduces it later): the two add up to the oracle's
    assert abs(sum(float(outs[r][3]) for r in range(world)) - float(scale.grad)) < 1e-4 * max(1.0, abs(float(scale.grad)))
