"""Golden fixture for the training step of the synchronisation module, from the UNMODIFIED reference.

    python tests/golden/make_golden_train.py

Takes the reference `Synchformer` built by `_ref_import.py`, loads the synthetic weights, puts `vproj` / `aproj` / `transformer`
in train mode and runs them on the committed golden features (`sync_b2s2.npz`: the frozen extractors' outputs) with
  (a) every dropout probability 0, and
  (b) the reference's `nn.Dropout` INSTANCES swapped for modules that multiply by the counter-based multipliers of
      `oracle/philox.py` (seed below) — the reference's code is untouched, only the stochastic source is made explicit,
then `F.cross_entropy(logits, targets).backward()` (train_utils.py:373-386 without the scaler).  Stored per case: logits, loss and
for each of the 50 trainable tensors its gradient's l2 norm, sum and a strided sample.
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

B, S, SEED_W, SEED_DROP, SAMPLE_STRIDE = 2, 2, 1337, 20240611, 1009


class _Mul(torch.nn.Module):
    def __init__(self, m):
        super().__init__()
        self.m = m

    def forward(self, x):
        return x * self.m


def grads_of(model, vfeats, afeats, targets):
    model.zero_grad(set_to_none=True)
    with torch.enable_grad():
        v, a = model.vproj(vfeats), model.aproj(afeats)                                           # sync_model.py:55-56
        Bb, Ss = v.shape[:2]
        logits = model.transformer(v.reshape(Bb, Ss * 8, 768), a.reshape(Bb, Ss * 6, 768))        # :59-66
        loss = model.compute_loss(logits, targets)
        loss.backward()
    out = {}
    for n, p in model.named_parameters():
        if n.split('.')[0] in ('vproj', 'aproj', 'transformer'):
            g = p.grad.detach().double().reshape(-1)
            out[n] = (np.array([float(g.norm()), float(g.sum())]), g[::SAMPLE_STRIDE].float().numpy())
    return logits.detach(), float(loss), out


def main():
    torch.manual_seed(0)
    cwd = os.getcwd()
    model = _ref_import.build_reference_model(n_segments=S)
    sd = synth.synthetic_state_dict(SEED_W, n_segments=S)
    model.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=True)
    os.chdir(cwd)
    gold = np.load(os.path.join(HERE, 'sync_b2s2.npz'))
    vfeats, afeats = torch.from_numpy(gold['vfeats']), torch.from_numpy(gold['afeats'])
    targets = torch.from_numpy(gold['targets'])
    T = 2 + 14 * S
    tr = model.transformer
    model.train()
    for n, p in model.named_parameters():                  # get_model freezes the extractors only (train_utils.py:199-204)
        p.requires_grad_(n.split('.')[0] in ('vproj', 'aproj', 'transformer'))
    store = {}

    # (a) dropout off
    drops = [m for m in tr.modules() if isinstance(m, torch.nn.Dropout)]
    saved_p = [m.p for m in drops]
    for m in drops:
        m.p = 0.0
    logits, loss, g = grads_of(model, vfeats, afeats, targets)
    print('p=0   loss', loss, 'logits[0,:4]', logits[0, :4].tolist())
    store['p0_logits'], store['p0_loss'] = logits.numpy(), np.float64(loss)
    for n, (stat, sample) in g.items():
        store['p0_stat/' + n], store['p0_sample/' + n] = stat, sample
    for m, p in zip(drops, saved_p):
        m.p = p
    assert (tr.drop.p, tr.blocks[0].attn.attn_drop.p, tr.blocks[0].attn.resid_drop.p, tr.blocks[0].mlp[3].p) == (0.1, 0.1, 0.1, 0.1)

    # (b) explicit multipliers in place of the nn.Dropout instances
    mult = O.train_multipliers(B, T, SEED_DROP, 0.1, 0.1, 0.1)
    tr.drop = _Mul(mult['embd'])
    for i, blk in enumerate(tr.blocks):
        blk.attn.attn_drop = _Mul(mult[f'attn{i}'])
        blk.attn.resid_drop = _Mul(mult[f'resid_attn{i}'])
        blk.mlp[3] = _Mul(mult[f'resid_mlp{i}'])
    logits, loss, g = grads_of(model, vfeats, afeats, targets)
    print('p=0.1 loss', loss, 'logits[0,:4]', logits[0, :4].tolist())
    store['drop_logits'], store['drop_loss'] = logits.numpy(), np.float64(loss)
    for n, (stat, sample) in g.items():
        store['drop_stat/' + n], store['drop_sample/' + n] = stat, sample

    np.savez_compressed(os.path.join(HERE, 'sync_train_b2s2.npz'), meta=np.array([B, S, SEED_W, SEED_DROP, SAMPLE_STRIDE]), **store)
    print('wrote', os.path.join(HERE, 'sync_train_b2s2.npz'), len(g), 'trainable tensors')


if __name__ == '__main__':
    main()
