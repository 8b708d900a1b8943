"""CPU: the oracle restatement against the committed outputs of the reference itself (tests/golden/make_golden.py)."""
import os

import numpy as np
import torch

from oracle import synchformer_oracle as O
from synchformer_b200 import synth


def rel_l2(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def _load(golden_dir):
    g = np.load(os.path.join(golden_dir, 'sync_b2s2.npz'))
    B, S, seed_w, seed_x, st_tok, st_d = [int(v) for v in g['meta']]
    sd = synth.synthetic_state_dict(seed_w, n_segments=S)
    return g, sd, B, S, seed_x, st_tok, st_d


def test_synthetic_weights_are_the_ones_the_golden_was_made_with(golden_dir):
    g, sd, *_ = _load(golden_dir)
    for name, (s, a) in zip(g['weight_checksum_names'], g['weight_checksums']):
        t = sd[str(name)].double()
        assert abs(float(t.sum()) - s) <= 1e-6 * max(1.0, abs(a)), name
        assert abs(float(t.abs().sum()) - a) <= 1e-6 * a, name


def test_mel_frontend_matches_reference_transforms(golden_dir):
    m = np.load(os.path.join(golden_dir, 'mel_b2s2.npz'))
    wave = synth.synthetic_waveform(2, 2, 0)
    assert abs(float(wave.double().sum()) - m['wave_checksum'][0]) < 1e-3
    mel = O.mel_frontend(wave).float().numpy()
    assert mel.shape == (2, 2, 128, 66)
    # the reference (torchaudio, fp32 FFT) is itself ~2e-4 away from the fp64 DFT on near-silent bins
    assert np.abs(mel - m['mel']).max() < 1e-3
    assert np.abs(mel - m['mel']).mean() < 5e-5
    # padded 66th frame: (0 + 4.2677393) / (2 * 4.5689974)
    np.testing.assert_allclose(mel[..., 65], 0.4670, atol=1e-4)


def test_forward_matches_reference_outputs(golden_dir):
    torch.set_grad_enabled(False)
    g, sd, B, S, seed_x, st_tok, st_d = _load(golden_dir)
    vis = synth.synthetic_video(B, S, seed_x)
    aud = torch.from_numpy(np.load(os.path.join(golden_dir, 'mel_b2s2.npz'))['mel']).unsqueeze(2)
    taps = {}
    loss, logits = O.forward(sd, vis, aud, torch.from_numpy(g['targets']), taps=taps)
    assert rel_l2(taps['vfeats'], g['vfeats']) < 1e-5
    assert rel_l2(taps['afeats'], g['afeats']) < 1e-5
    for k in ('v_embed', 'v_block0', 'v_block11', 'a_embed', 'a_last_hidden'):
        assert rel_l2(taps[k][:, ::st_tok, ::st_d], g[k]) < 1e-5, k
    assert np.abs(logits.numpy() - g['logits']).max() < 1e-4
    assert (logits.argmax(-1).numpy() == g['logits'].argmax(-1)).all()
    assert abs(float(loss) - float(g['loss'])) < 1e-4
    # inputs must matter (SURVEY.md §0 trap 1): the two clips' features differ
    assert rel_l2(taps['vfeats'][0], taps['vfeats'][1]) > 1e-2


def test_cls_row_shortcut_is_exact():
    """The product evaluates the CLS aggregators for the CLS row only; the oracle evaluates them densely like the reference."""
    torch.manual_seed(0)
    sd = {k: v for k, v in synth.synthetic_state_dict(7, 2).items() if k.startswith('vfeat_extractor.spatial_attn_agg.')}
    p = 'vfeat_extractor.spatial_attn_agg.'
    x = torch.randn(3, 196, 768)
    dense = O.cls_aggregator(sd, p, x)
    xx = torch.cat([sd[p + 'cls_token'].expand(3, 1, 768), x], 1)
    y = O._ln(xx, sd[p + 'norm1.weight'], sd[p + 'norm1.bias'], 1e-6)
    w, b = sd[p + 'self_attn.in_proj_weight'], sd[p + 'self_attn.in_proj_bias']
    q = (y[:, :1] @ w[:768].T + b[:768]).view(3, 1, 12, 64).transpose(1, 2) * 0.125
    k = (y @ w[768:1536].T + b[768:1536]).view(3, 197, 12, 64).transpose(1, 2)
    v = (y @ w[1536:].T + b[1536:]).view(3, 197, 12, 64).transpose(1, 2)
    a = (torch.softmax(q @ k.transpose(-1, -2), -1) @ v).transpose(1, 2).reshape(3, 768)
    y0 = xx[:, 0] + a @ sd[p + 'self_attn.out_proj.weight'].T + sd[p + 'self_attn.out_proj.bias']
    h = O._gelu(O._ln(y0, sd[p + 'norm2.weight'], sd[p + 'norm2.bias'], 1e-6) @ sd[p + 'linear1.weight'].T + sd[p + 'linear1.bias'])
    row = y0 + h @ sd[p + 'linear2.weight'].T + sd[p + 'linear2.bias']
    assert (dense - row).abs().max() < 2e-5
