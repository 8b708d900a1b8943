"""Golden fixture for the stage-I training step of the two encoders, from the UNMODIFIED reference.

    python tests/golden/make_golden_encoders_train.py

The reference `MotionFormer` and `AST` (the instances inside the reference `Synchformer` built by `_ref_import.py`, synthetic weights
loaded with strict=True) are put in TRAIN mode; the `DropPath` instances of the 12 Motionformer blocks (vit_helper.py:356) are swapped
for modules that multiply by the counter-based per-segment multipliers of `oracle/philox.py` (seed below; rates 0.2 i / 11) - the
reference's code is untouched, only the stochastic source is made explicit.  Loss = the stage-I contrastive loss
(`AVCLIP.compute_loss`, open_clip/model.py:507-527: time-mean pooling, L2-normalise, symmetric cross-entropy, logit scale 0.07) restated
on the two feature tensors (importing the reference `AVCLIP` class needs `ftfy`, which is not installed).  Stored: loss, features, and
for each of the 448 participating tensors its gradient's l2 norm, sum and a strided sample.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, HERE)

from oracle import synchformer_oracle as O  # noqa: E402
from synchformer_b200 import synth  # noqa: E402
import _ref_import  # noqa: E402

B, S, SEED_W, SEED_X, SEED_DROP, SAMPLE_STRIDE = 1, 2, 1337, 0, 7, 4099


class _Mul(torch.nn.Module):
    def __init__(self, m):
        super().__init__()
        self.m = m

    def forward(self, x):
        return x * self.m


def main():
    torch.manual_seed(0)
    cwd = os.getcwd()
    model = _ref_import.build_reference_model(n_segments=S)
    sd = synth.synthetic_state_dict(SEED_W, n_segments=S)
    model.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=True)
    os.chdir(cwd)
    vis = synth.synthetic_video(B, S, SEED_X)
    aud = O.mel_frontend(synth.synthetic_waveform(B, S, SEED_X)).float().unsqueeze(2)
    vfe, afe = model.vfeat_extractor, model.afeat_extractor
    for n, p in model.named_parameters():
        p.requires_grad_(n.split('.')[0] in ('vfeat_extractor', 'afeat_extractor') and '.patch_embed.proj.' not in n)
    vfe.train()
    afe.train()
    mult = O.drop_path_multipliers(B * S, SEED_DROP)
    assert abs(vfe.blocks[11].drop_path.drop_prob - 0.2) < 1e-6 and abs(vfe.blocks[5].drop_path.drop_prob - 0.2 * 5 / 11) < 1e-6
    assert not hasattr(vfe.blocks[0].drop_path, 'drop_prob')              # rate 0 -> nn.Identity (vit_helper.py:356)

    class _SpaceThenMlp(torch.nn.Module):
        """block.drop_path is called twice per forward: on the space-attention branch, then on the MLP branch (vit_helper.py:371,375)"""

        def __init__(self, ms, mm):
            super().__init__()
            self.m, self.k = (ms, mm), 0

        def forward(self, x):
            m = self.m[self.k % 2]
            self.k += 1
            return x * m

    for i, blk in enumerate(vfe.blocks):
        blk.drop_path = _SpaceThenMlp(*mult[i])
    with torch.enable_grad():
        vfeat = model.extract_vfeats(vis, for_loop=False)
        afeat = model.extract_afeats(aud, for_loop=False)
        loss = O.contrastive_loss(vfeat, afeat, 0.07)
        loss.backward()
    store = dict(loss=np.float64(float(loss)), vfeats=vfeat.detach().numpy(), afeats=afeat.detach().numpy())
    n_t = 0
    for n, p in model.named_parameters():
        if p.requires_grad:
            g = p.grad.detach().double().reshape(-1)
            store['stat/' + n] = np.array([float(g.norm()), float(g.sum())])
            store['sample/' + n] = g[::SAMPLE_STRIDE].float().numpy()
            n_t += 1
    print('loss', float(loss), 'tensors', n_t)
    np.savez_compressed(os.path.join(HERE, 'encoders_train_b1s2.npz'), meta=np.array([B, S, SEED_W, SEED_X, SEED_DROP, SAMPLE_STRIDE]), **store)
    print('wrote', os.path.join(HERE, 'encoders_train_b1s2.npz'))


if __name__ == '__main__':
    main()
