"""CPU, build container only: re-run the LIVE reference (read-only /root/reference) against the oracle on fresh inputs.
Skipped wherever the reference is not mounted (e.g. the GPU box)."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))
import _ref_import  # noqa: E402

pytestmark = pytest.mark.skipif(not _ref_import.reference_available(), reason='reference checkout not mounted')


def test_oracle_equals_live_reference_on_new_inputs():
    from oracle import synchformer_oracle as O
    from synchformer_b200 import synth
    torch.set_grad_enabled(False)
    cwd = os.getcwd()
    try:
        model = _ref_import.build_reference_model(n_segments=1)
    finally:
        os.chdir(cwd)
    sd = synth.synthetic_state_dict(4242, n_segments=1)
    model.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=True)
    vis = synth.synthetic_video(2, 1, seed=5)
    aud = O.mel_frontend(synth.synthetic_waveform(2, 1, seed=5)).float().unsqueeze(2)
    _, ref_logits = model(vis, aud)
    ref_v = model.extract_vfeats(vis, for_loop=False)
    taps = {}
    _, logits = O.forward(sd, vis, aud, taps=taps)
    assert np.abs(logits.numpy() - ref_logits.numpy()).max() < 1e-4
    assert float((taps['vfeats'] - ref_v).norm() / ref_v.norm()) < 1e-5
