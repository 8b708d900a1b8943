"""Generate the committed golden fixtures from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py

Imports the reference from /root/reference through `_ref_import.py` (in-memory stubs for omegaconf / timm /
transformers-4.27 APIs; no reference file is copied), loads `synchformer_b200.synth.synthetic_state_dict`
into it with `load_state_dict(strict=True)`, runs `Synchformer.forward` in eval / fp32 on the synthetic
inputs and stores outputs + a few strided intermediate taps in `tests/golden/sync_b2s2.npz`.  The mel
front-end golden (`mel_b2s2.npz`) is produced with the reference transform classes
(dataset/transforms.py:815-871, i.e. torchaudio).  Weight checksums are stored so a torch RNG change is detected
instead of silently comparing different models.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, HERE)

from synchformer_b200 import synth  # noqa: E402
import _ref_import  # noqa: E402

B, S, SEED_W, SEED_X = 2, 2, 1337, 0
TAP_STRIDE_TOK, TAP_STRIDE_D = 97, 7


def sub(t: torch.Tensor) -> np.ndarray:
    return t[:, ::TAP_STRIDE_TOK, ::TAP_STRIDE_D].contiguous().float().numpy()


def checksums(sd):
    names = ['vfeat_extractor.patch_embed_3d.proj.weight', 'vfeat_extractor.blocks.7.timeattn.qkv.weight',
             'afeat_extractor.ast.encoder.layer.3.output.dense.weight', 'transformer.pos_emb_cfg.pos_emb',
             'transformer.off_head.weight']
    return names, np.array([[float(sd[n].double().sum()), float(sd[n].double().abs().sum())] for n in names])


def main():
    torch.manual_seed(0)
    torch.set_grad_enabled(False)
    cwd = os.getcwd()
    model = _ref_import.build_reference_model(n_segments=S)
    sd = synth.synthetic_state_dict(SEED_W, n_segments=S)
    missing = model.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=True)
    print('load_state_dict:', missing)
    assert len(model.state_dict()) == 513

    vis = synth.synthetic_video(B, S, SEED_X)
    wave = synth.synthetic_waveform(B, S, SEED_X)

    # ---- mel front-end through the reference transform classes (dataset/transforms.py) -------------------
    # dataset.transforms imports torchvision/torchaudio; only the audio tail is exercised here.
    import importlib
    tr = importlib.import_module('dataset.transforms')
    chain = [tr.AudioMelSpectrogram(sample_rate=16000, win_length=400, hop_length=160, n_fft=1024, n_mels=128),
             tr.AudioLog(), tr.PadOrTruncate(max_spec_t=66), tr.AudioNormalizeAST(mean=-4.2677393, std=4.5689974)]
    mels = []
    for b in range(B):
        item = {'audio': wave[b].clone(), 'meta': {'audio': {}}}
        for t in chain:
            item = t(item)
        mels.append(item['audio'])
    mel = torch.stack(mels)                                   # (B, S, 128, 66)
    aud = mel.unsqueeze(2)                                    # PermuteStreams 'S F T -> S 1 F T'

    taps = {}
    vfe, afe = model.vfeat_extractor, model.afeat_extractor
    hooks = [
        vfe.pos_drop.register_forward_hook(lambda m, i, o: taps.__setitem__('v_embed', o.detach())),
        vfe.blocks[0].register_forward_hook(lambda m, i, o: taps.__setitem__('v_block0', o.detach())),
        vfe.blocks[11].register_forward_hook(lambda m, i, o: taps.__setitem__('v_block11', o.detach())),
        afe.ast.embeddings.register_forward_hook(lambda m, i, o: taps.__setitem__('a_embed', o.detach())),
        afe.ast.layernorm.register_forward_hook(lambda m, i, o: taps.__setitem__('a_last_hidden', o.detach())),
    ]
    vfeats = model.extract_vfeats(vis, for_loop=False)
    afeats = model.extract_afeats(aud, for_loop=False)
    for h in hooks:
        h.remove()
    targets = torch.tensor([3, 17])
    loss, logits = model(vis, aud, targets)
    print('logits', logits)
    print('loss', float(loss), 'argmax', logits.argmax(-1).tolist())
    d = (vfeats[0] - vfeats[1]).norm() / vfeats[0].norm()
    print('relative difference between the two clips\' visual features:', float(d))
    assert d > 1e-2, 'inputs must influence the features'

    # noise floor of the reference itself: its bf16-autocast path against its fp32 path on the same weights / inputs
    # (SURVEY.md §8d calibrates the parity gates to this; stored so the tests can state the gate next to the floor)
    with torch.autocast('cpu', dtype=torch.bfloat16):
        vf16 = model.extract_vfeats(vis, for_loop=False).float()
        af16 = model.extract_afeats(aud, for_loop=False).float()
        lg16 = model(vis, aud)[1].float()
    rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
    ref_bf16_noise = np.array([rel(vf16, vfeats), rel(af16, afeats), float((lg16 - logits).abs().max()), rel(lg16, logits)])
    print('reference bf16-autocast vs fp32 [vfeats rel, afeats rel, logits max-abs, logits rel]:', ref_bf16_noise)

    names, cs = checksums(sd)
    os.chdir(cwd)
    np.savez_compressed(
        os.path.join(HERE, 'sync_b2s2.npz'),
        vfeats=vfeats.numpy(), afeats=afeats.numpy(), logits=logits.numpy(), loss=np.float32(loss),
        targets=targets.numpy(),
        **{k: sub(v) for k, v in taps.items()},
        weight_checksum_names=np.array(names), weight_checksums=cs, ref_bf16_noise=ref_bf16_noise,
        meta=np.array([B, S, SEED_W, SEED_X, TAP_STRIDE_TOK, TAP_STRIDE_D]),
    )
    np.savez_compressed(os.path.join(HERE, 'mel_b2s2.npz'), mel=mel.numpy(),
                        wave_checksum=np.array([float(wave.double().sum()), float(wave.double().abs().sum())]))
    print('wrote fixtures to', HERE)


if __name__ == '__main__':
    main()
