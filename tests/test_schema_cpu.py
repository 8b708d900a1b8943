"""CPU: the drop-in surface — constructor from the sync.yaml config tree, state-dict schema, pos-emb trimming, error behaviour."""
import pytest
import torch

from synchformer_b200 import model as M
from synchformer_b200 import schema, synth


def test_state_dict_schema_is_the_references():
    m = M.build_synchformer(n_segments=14)
    sd = m.state_dict()
    ref = schema.state_dict_schema(14)
    assert len(sd) == 513
    assert list(sd.keys()) == list(ref.keys())
    assert all(tuple(sd[k].shape) == ref[k] for k in ref)
    assert sum(v.numel() for v in sd.values()) == 237_460_245   # parameter census of the reference
    assert all(v.dtype == torch.float32 for v in sd.values())


def test_constructor_accepts_reference_config_tree():
    cfg = M.sync_yaml_model_config(n_segments=14)
    cfg = {k: dict(v) for k, v in cfg.items()}
    for k in ('afeat_extractor', 'vfeat_extractor'):
        cfg[k].pop('is_trainable')                          # get_model pops nothing: instantiate ignores extra keys next to target/params
    m = M.Synchformer(**cfg)
    assert isinstance(m.vfeat_extractor, torch.nn.Module) and isinstance(m.afeat_extractor, torch.nn.Module)
    assert m.transformer.pos_emb_cfg.pos_emb.shape == (1, 198, 768)
    # scripts/train_utils.py:199-204 freezes the extractors through .parameters()
    for p in m.vfeat_extractor.parameters():
        p.requires_grad = False
    assert not m.vfeat_extractor.patch_embed.proj.weight.requires_grad
    assert m.__class__.__name__ == 'Synchformer'


def test_load_state_dict_strict_and_posemb_trimming():
    m = M.build_synchformer(n_segments=2)
    sd = synth.synthetic_state_dict(1, n_segments=2)
    res = m.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    long_sd = synth.synthetic_state_dict(1, n_segments=14)
    m.load_state_dict(long_sd)                                # longer table is trimmed (sync_model.py:109-111)
    assert m.transformer.pos_emb_cfg.pos_emb.shape[1] == 30
    assert long_sd['transformer.pos_emb_cfg.pos_emb'].shape[1] == 198     # caller's dict untouched
    with pytest.raises(ValueError):
        M.build_synchformer(n_segments=14).load_state_dict(sd)           # shorter table is an error (:112-113)
    bad = dict(sd)
    bad.pop('vproj.weight')
    with pytest.raises(RuntimeError):
        m.load_state_dict(bad, strict=True)


def test_unsupported_configurations_fail_loudly():
    with pytest.raises(NotImplementedError):
        M.MotionFormer(extract_features=False)
    with pytest.raises(NotImplementedError):
        M.AST(extract_features=True, factorize_freq_time=True, agg_freq_module='AveragePooling', agg_time_module='torch.nn.Identity',
              add_global_repr=False)
    with pytest.raises(NotImplementedError):
        M.instantiate_from_config({'target': 'model.modules.feat_extractors.visual.s3d.S3DVisualFeatures', 'params': {}})
    with pytest.raises(KeyError):
        M.instantiate_from_config({'params': {}})


def test_no_cpu_fallback():
    from synchformer_b200._lib import SfbError
    m = M.build_synchformer(n_segments=1)
    with pytest.raises(SfbError):
        m(torch.zeros(1, 1, 16, 3, 224, 224), torch.zeros(1, 1, 1, 128, 66))
    with pytest.raises(NotImplementedError):
        m.extract_vfeats(torch.zeros(1, 1, 16, 3, 224, 224), vis_mask=torch.ones(1))
    assert m.compute_loss(torch.zeros(2, 21), None) is None
    with pytest.raises(NotImplementedError):
        m.compute_loss(torch.zeros(2, 21), torch.zeros(2, dtype=torch.long), loss_fn='focal')


def test_syncability_head_variant():
    cfg = M.sync_yaml_model_config(n_segments=13, transformer_target='model.sync_model.GlobalTransformerWithSyncabilityHead')
    cfg = {k: {kk: vv for kk, vv in v.items() if kk != 'is_trainable'} for k, v in cfg.items()}
    m = M.Synchformer(**cfg)
    assert m.transformer.pos_emb_cfg.pos_emb.shape == (1, 184, 768)     # configs/ft_synchability.yaml:55
    assert m.transformer.sync_head.weight.shape == (2, 768)
    assert 'transformer.off_head.weight' not in m.state_dict()
