"""CPU: host orchestration of the fused-LayerNorm schedule of the Motionformer blocks (model.MotionFormer._encode_chunk with
`fuse_layernorm`): weight folding (gamma into W, beta into the bias, column sums of the folded bf16 weights), the hand-over of the bf16
stream copy and the partial row statistics from every residual GEMM to the next qkv / fc1 GEMM, and the unfused schedule kept for A/B.
Every kernel is replaced by the fp32 restatement of its C-ABI contract (tests/fake_ops.py), so the result must equal the fp32 oracle of the
reference forward (vit_helper.py:364-376, motionformer.py:229-252) to round-off - any wrong operand, slice or statistic shows up as O(1)."""
import pytest
import torch

from oracle import synchformer_oracle as O
from synchformer_b200 import model as M, synth

import fake_ops


@pytest.mark.parametrize('fused', [1, 2, 0])        # 1: all three norms fused, 2: the two in front of the qkv GEMMs, 0: LayerNorm launches
def test_motionformer_schedules_equal_the_oracle(monkeypatch, fused):
    fake_ops.install(monkeypatch, round_bf16=False, names=fake_ops.ALL + fake_ops.ENCODER_FWD + ('empty_bf16',))
    sd = synth.synthetic_state_dict(3, n_segments=1)
    model = M.build_synchformer(n_segments=1, state_dict=sd)
    model.vfeat_extractor.fuse_layernorm = fused
    vis = synth.synthetic_video(1, 1, 5)
    taps = {}
    model.vfeat_extractor._taps = taps
    with torch.no_grad():
        vf = model.extract_vfeats(vis)
    ref_taps = {}
    ref = O.extract_vfeats(sd, vis, taps=ref_taps)
    for k in ('v_embed', 'v_block0', 'v_block11'):
        err = float((taps[k] - ref_taps[k]).norm() / ref_taps[k].norm())
        assert err < 2e-5, (k, err)
    assert float((vf - ref).norm() / ref.norm()) < 5e-5


def test_folded_weights_are_what_the_header_says(monkeypatch):
    fake_ops.install(monkeypatch, round_bf16=False, names=('require_cuda', 'cast_bf16'))
    monkeypatch.setattr(fake_ops, 'REAL_DTYPES', True)
    sd = synth.synthetic_state_dict(4, n_segments=1)
    model = M.build_synchformer(n_segments=1, state_dict=sd)
    P, W = model.vfeat_extractor.weights()
    b = 'blocks.7.'
    for norm, lin in (('norm3', 'timeattn.qkv'), ('norm1', 'attn.qkv'), ('norm2', 'mlp.fc1')):
        w0, b0, g, be = P[b + lin + '.weight'], P[b + lin + '.bias'], P[b + norm + '.weight'], P[b + norm + '.bias']
        wf = W[b + lin + '.fold_w']
        assert wf.dtype == torch.bfloat16 and torch.equal(wf, (w0 * g).to(torch.bfloat16))
        assert torch.allclose(W[b + lin + '.fold_cs'], wf.float().sum(1), rtol=1e-6, atol=1e-6)
        assert torch.allclose(W[b + lin + '.fold_b'], b0 + w0 @ be, rtol=1e-5, atol=1e-6)
