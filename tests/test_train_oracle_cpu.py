"""CPU: the training-step oracle against the golden fixture made from the live reference, the Philox restatement
against the published known-answer vectors, and the host orchestration of `synchformer_b200/train.py` (run on CPU stand-ins for the
kernels, tests/fake_ops.py) against the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import philox
from oracle import synchformer_oracle as O
from synchformer_b200 import model as M, synth

import fake_ops
import train_gates

FLOOR = 1e-6          # absolute l2 floor below which a gradient tensor counts as zero (typical gradient norms here: 1e-3 .. 1)
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, 'golden')


@pytest.fixture(autouse=True)
def _grad_enabled():
    """other test modules import the reference, which switches autograd off process-wide"""
    with torch.enable_grad():
        yield


def test_philox4x32_10_known_answers():
    """Random123 kat_vectors for philox4x32-10: (counter, key) -> output."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = philox.philox4x32_10(*ctr, *key)
        assert tuple(int(g) for g in got) == want


def test_dropout_multiplier_properties():
    m = philox.dropout_multiplier((257, 768), 0.1, 99, 4)
    assert set(np.unique(m).tolist()) == {0.0, float(np.float32(1.0) / (np.float32(1.0) - np.float32(0.1)))}
    assert abs(float((m == 0).mean()) - 0.1) < 5e-3
    assert np.array_equal(m, philox.dropout_multiplier((257, 768), 0.1, 99, 4))            # pure function of (seed, site, index)
    assert not np.array_equal(m, philox.dropout_multiplier((257, 768), 0.1, 99, 5))        # sites are independent streams
    assert not np.array_equal(m, philox.dropout_multiplier((257, 768), 0.1, 100, 4))
    assert np.array_equal(m.reshape(-1)[:1001], philox.dropout_multiplier((1001,), 0.1, 99, 4))   # prefix-stable in the linear index
    assert np.all(philox.dropout_multiplier((8, 8), 0.0, 1, 0) == 1.0)
    assert philox.threshold(0.1) == int(np.floor(float(np.float32(0.1)) * 2 ** 32))


@pytest.fixture(scope='module')
def gold():
    return np.load(os.path.join(GOLD, 'sync_train_b2s2.npz')), np.load(os.path.join(GOLD, 'sync_b2s2.npz'))


@pytest.fixture(scope='module')
def setup(gold):
    g, f = gold
    B, S, seed_w, seed_drop, stride = (int(x) for x in g['meta'])
    sd = synth.synthetic_state_dict(seed_w, n_segments=S)
    return dict(B=B, S=S, T=2 + 14 * S, sd=sd, seed=seed_drop, stride=stride, vf=torch.from_numpy(f['vfeats']), af=torch.from_numpy(f['afeats']),
                targets=torch.from_numpy(f['targets']))


@pytest.mark.parametrize('case', ['p0', 'drop'])
def test_oracle_train_step_matches_reference_golden(gold, setup, case):
    """Oracle autograd vs the reference's own backward (tests/golden/make_golden_train.py), fp32 both: loss, logits, and for each of
    the 63 trainable tensors the gradient norm, sum and a strided sample."""
    g, _ = gold
    s = setup
    mult = None if case == 'p0' else O.train_multipliers(s['B'], s['T'], s['seed'])
    loss, logits, grads = O.sync_train_grads(s['sd'], s['vf'], s['af'], s['targets'], mult)
    assert abs(float(loss) - float(g[case + '_loss'])) < 1e-5
    assert np.abs(logits.numpy() - g[case + '_logits']).max() < 2e-5
    assert len(grads) == 63
    for n, gr in grads.items():
        flat = gr.double().reshape(-1)
        stat, sample = g[f'{case}_stat/{n}'], g[f'{case}_sample/{n}']
        # (the key bias has a mathematically zero gradient - softmax is invariant to a per-query shift - so it is pure round-off: FLOOR)
        assert abs(float(flat.norm()) - stat[0]) <= 2e-4 * stat[0] + FLOOR, n
        err = np.abs(flat[::s['stride']].float().numpy() - sample).max()
        assert err <= 2e-4 * np.abs(sample).max() + FLOOR, (n, err)


def _train_model(setup, p_drop: float):
    s = setup
    cfg = M.sync_yaml_model_config(s['S'])
    for k in ('embd_pdrop', 'resid_pdrop', 'attn_pdrop'):
        cfg['transformer']['params'][k] = p_drop
    cfg = {k: {kk: vv for kk, vv in v.items() if kk != 'is_trainable'} for k, v in cfg.items()}
    model = M.Synchformer(**cfg)
    model.load_state_dict(s['sd'], strict=True)
    model.train()
    for ext in (model.vfeat_extractor, model.afeat_extractor):       # train_utils.py:199-204, 330-342
        ext.requires_grad_(False)
        ext.eval()
    return model


def _host_grads(model, setup, monkeypatch, seed):
    s = setup
    from synchformer_b200 import train
    monkeypatch.setattr(train, 'draw_seed', lambda: seed)
    model.zero_grad(set_to_none=True)
    v, a = model.project(s['vf'], s['af'])
    logits = model.transformer(v, a)
    loss = model.compute_loss(logits, s['targets'])
    loss.backward()
    grads = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
    return loss.detach(), logits.detach(), grads


@pytest.mark.parametrize('p_drop', [0.0, 0.1])
def test_host_orchestration_matches_oracle_exactly_in_fp32(setup, monkeypatch, p_drop):
    """train.py on fp32 CPU stand-ins for the kernels == oracle autograd: every operand / transpose / slice / gather / site id is right."""
    s = setup
    fake_ops.install(monkeypatch, round_bf16=False)
    model = _train_model(s, p_drop)
    loss, logits, grads = _host_grads(model, s, monkeypatch, s['seed'])
    mult = None if p_drop == 0 else O.train_multipliers(s['B'], s['T'], s['seed'], p_drop, p_drop, p_drop)
    rloss, rlogits, rgrads = O.sync_train_grads(s['sd'], s['vf'], s['af'], s['targets'], mult)
    assert abs(float(loss) - float(rloss)) < 1e-5
    assert (logits - rlogits).abs().max() < 5e-5
    assert set(grads) == set(rgrads), set(rgrads) ^ set(grads)
    for n, r in rgrads.items():
        assert grads[n].shape == r.shape, n
        err, ref = float((grads[n].double() - r.double()).norm()), float(r.double().norm())
        assert err <= 2e-4 * ref + FLOOR, (n, err, ref)


def test_host_orchestration_bf16_noise_estimate(setup, monkeypatch):
    """Same with the stand-ins rounding to bf16 wherever the kernels do: the gate used on the GPU (tests/test_train_gpu.py: per-tensor
    gradient rel-L2 <= 3e-2 against the fp32 oracle) has to hold here with margin."""
    s = setup
    fake_ops.install(monkeypatch, round_bf16=True)
    model = _train_model(s, 0.1)
    loss, logits, grads = _host_grads(model, s, monkeypatch, s['seed'])
    mult = O.train_multipliers(s['B'], s['T'], s['seed'])
    rloss, rlogits, rgrads = O.sync_train_grads(s['sd'], s['vf'], s['af'], s['targets'], mult)
    worst = train_gates.check_grads(grads, rgrads)
    print('worst gradient error / allowance with bf16 rounding:', worst, 'logits max-abs', float((logits - rlogits).abs().max()))
    assert worst < 0.6                                              # margin: the GPU gate must not sit at the edge of the noise
    assert (logits - rlogits).abs().max() < 2e-2


def test_eval_mode_without_grad_does_not_take_the_training_path(setup, monkeypatch):
    s = setup
    fake_ops.install(monkeypatch)
    from synchformer_b200 import train
    called = []
    monkeypatch.setattr(train, 'sync_transformer', lambda *a, **k: called.append(1))
    monkeypatch.setattr(train, 'linear', lambda *a, **k: called.append(2))
    model = _train_model(s, 0.1).eval()
    with torch.no_grad():
        try:
            v, a = model.project(s['vf'], s['af'])
        except Exception:
            pass
    assert not called


def test_tok_pdrop_is_refused(setup, monkeypatch):
    fake_ops.install(monkeypatch)
    model = _train_model(setup, 0.1)
    model.transformer.tok_pdrop = 0.2
    with pytest.raises(NotImplementedError, match='tok_pdrop'):
        model.transformer(torch.zeros(1, 8 * setup['S'], 768), torch.zeros(1, 6 * setup['S'], 768))
