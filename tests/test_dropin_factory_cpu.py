"""CPU: the drop-in loop closed against the reference's OWN factory and its REAL YAML files (VERDICT r1 "missing" #6).

`utils.utils.instantiate_from_config` (utils/utils.py:78-88) of the unmodified reference builds this package's classes from
configs/sync.yaml, configs/ft_synchability.yaml and configs/segment_avclip.yaml with ONLY `model.target` overridden - the containers it
passes are the harness' attribute-dict / list-like config nodes, not plain dicts.  Then what `get_model` does next
(scripts/train_utils.py:199-204: freeze by `is_trainable`), a strict state-dict exchange with a reference-built model in both directions,
and the stage-I `ckpt_path` initialisation of the extractors (motionformer.py:156-173, ast.py:113-132).

Needs the reference sources (/root/reference here, or the staged baseline/_ref): skipped where neither exists."""
import logging
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'golden'))
import _ref_import  # noqa: E402

if not _ref_import.reference_available():
    _staged = os.path.join(os.path.dirname(HERE), 'baseline', '_ref')
    if os.path.isdir(os.path.join(_staged, 'model')):
        _ref_import.REF_ROOT = _staged
pytestmark = pytest.mark.skipif(not _ref_import.reference_available(), reason='reference sources not present')


@pytest.fixture(scope='module')
def ref():
    cwd = os.getcwd()
    sync_model, _ = _ref_import.import_reference()
    import importlib
    utils = importlib.import_module('utils.utils')
    from omegaconf import OmegaConf
    yield sync_model, utils, OmegaConf
    os.chdir(cwd)
    torch.set_grad_enabled(True)


def _load(OmegaConf, name, n_segments=None):
    cfg = OmegaConf.load(os.path.join(_ref_import.REF_ROOT, 'configs', name))
    if n_segments is not None:
        cfg.model.params.transformer.params.pos_emb_cfg.params.block_shape = [2 + 14 * n_segments]
    return cfg


def test_reference_factory_builds_this_class_from_sync_yaml(ref):
    sync_model, utils, OmegaConf = ref
    cfg = _load(OmegaConf, 'sync.yaml')
    assert cfg.model.target == 'model.sync_model.Synchformer'
    cfg.model.target = 'synchformer_b200.model.Synchformer'                # the ONE key INTEGRATION.md changes
    model = utils.instantiate_from_config(cfg.model)                       # the reference's own factory, the harness' own containers
    from synchformer_b200 import model as M
    assert type(model) is M.Synchformer and type(model.vfeat_extractor) is M.MotionFormer and type(model.afeat_extractor) is M.AST
    assert type(model.transformer) is M.GlobalTransformer and isinstance(model.vproj, torch.nn.Linear)
    assert model.transformer.pos_emb_cfg.pos_emb.shape == (1, 198, 768) and model.transformer.off_head.weight.shape == (21, 768)
    # scripts/train_utils.py:199-204
    if cfg.model.params.vfeat_extractor.is_trainable is False:
        for params in model.vfeat_extractor.parameters():
            params.requires_grad = False
    if cfg.model.params.afeat_extractor.is_trainable is False:
        for params in model.afeat_extractor.parameters():
            params.requires_grad = False
    trainable = sum(p.numel() for p in model.parameters() if p.requires_grad)
    assert trainable == 22_619_157                                           # stage-II trainable parameters
    # the optimiser is built over ALL parameters (train_utils.py:225) and DDP needs >= 1 grad-requiring parameter
    assert sum(p.numel() for p in model.parameters()) == 237_460_245
    # strict state-dict exchange with a model the reference builds from the same file, both directions (checkpoints: logger.py:146, example.py:134)
    cfg_ref = _load(OmegaConf, 'sync.yaml')
    ref_model = utils.instantiate_from_config(cfg_ref.model)
    assert type(ref_model) is sync_model.Synchformer
    res = model.load_state_dict(ref_model.state_dict(), strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    res = ref_model.load_state_dict(model.state_dict(), strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert list(model.state_dict().keys()) == list(ref_model.state_dict().keys())
    assert model.__class__.__name__ == ref_model.__class__.__name__ == 'Synchformer'       # stored in checkpoints (logger.py:149)


def test_reference_factory_builds_the_syncability_and_avclip_variants(ref):
    _, utils, OmegaConf = ref
    from synchformer_b200 import avclip, model as M
    cfg = _load(OmegaConf, 'ft_synchability.yaml')
    cfg.model.target = 'synchformer_b200.model.Synchformer'
    model = utils.instantiate_from_config(cfg.model)
    assert type(model.transformer) is M.GlobalTransformerWithSyncabilityHead
    assert model.transformer.pos_emb_cfg.pos_emb.shape == (1, 184, 768) and model.transformer.sync_head.weight.shape == (2, 768)

    cfg = _load(OmegaConf, 'segment_avclip.yaml')
    assert cfg.model.target.endswith('AVCLIP')
    cfg.model.target = 'synchformer_b200.avclip.AVCLIP'
    for k in ('afeat_extractor', 'vfeat_extractor'):                     # the public pre-trained inits are downloads; random init here
        cfg.model.params[k].params.ckpt_path = None
    m = utils.instantiate_from_config(cfg.model)
    assert type(m) is avclip.AVCLIP and m.v_encoder.time_pool and m.a_encoder.time_pool
    assert float(m.logit_scale) == pytest.approx(cfg.model.params.init_scale)


def test_stage1_checkpoint_initialises_the_extractors(ref, tmp_path, caplog):
    """scripts/sbatch_train_sync.sh:74-75 passes the stage-I checkpoint as `ckpt_path` of both extractors; a reference-built AVCLIP-style
    state dict (keys `module.v_encoder.*` / `a_encoder.*`) must land in the towers, everything else of the file is ignored."""
    sync_model, utils, OmegaConf = ref
    from synchformer_b200 import model as M
    ref_model = utils.instantiate_from_config(_load(OmegaConf, 'sync.yaml').model)
    g = torch.Generator().manual_seed(3)
    state = {}
    for k, v in ref_model.vfeat_extractor.state_dict().items():
        state['module.v_encoder.' + k] = torch.randn(v.shape, generator=g) * 0.05          # DDP-wrapped naming
    for k, v in ref_model.afeat_extractor.state_dict().items():
        state['a_encoder.' + k] = torch.randn(v.shape, generator=g) * 0.05                 # plain naming
    state['module.logit_scale'] = torch.tensor(0.07)
    path = str(tmp_path / 'stage1.pt')
    torch.save({'state_dict': state, 'epoch': 3}, path)
    cfg = _load(OmegaConf, 'sync.yaml')
    cfg.model.target = 'synchformer_b200.model.Synchformer'
    cfg.model.params.vfeat_extractor.params.ckpt_path = path
    cfg.model.params.afeat_extractor.params.ckpt_path = path
    with caplog.at_level(logging.INFO):
        model = utils.instantiate_from_config(cfg.model)
    for k, v in model.vfeat_extractor.state_dict().items():
        assert torch.equal(v, state['module.v_encoder.' + k]), k
    for k, v in model.afeat_extractor.state_dict().items():
        assert torch.equal(v, state['a_encoder.' + k]), k
    assert not model.vfeat_extractor.patch_embed.proj.weight.requires_grad                   # motionformer.py:177 still holds after the load
    assert 'failed' not in caplog.text
    # same call on the reference class gives the same towers
    cfg_ref = _load(OmegaConf, 'sync.yaml')
    cfg_ref.model.params.vfeat_extractor.params.ckpt_path = path
    cfg_ref.model.params.afeat_extractor.params.ckpt_path = path
    ref_loaded = utils.instantiate_from_config(cfg_ref.model)
    for (k, a), (k2, b) in zip(model.vfeat_extractor.state_dict().items(), ref_loaded.vfeat_extractor.state_dict().items()):
        assert k == k2 and torch.equal(a, b), k
    # error behaviour: a missing file raises like utils/utils.py:57-58; a downloadable init is refused loudly
    with pytest.raises(ValueError, match='Cant find the checkpoint file'):
        M.MotionFormer(extract_features=True, ckpt_path=str(tmp_path / 'nope.pt'), factorize_space_time=True, agg_space_module='TransformerEncoderLayer',
                       agg_time_module='torch.nn.Identity', add_global_repr=False)
    with pytest.raises(NotImplementedError):
        M.AST(extract_features=True, ckpt_path='MIT/ast-finetuned-audioset-10-10-0.4593', max_spec_t=66, factorize_freq_time=True,
              agg_freq_module='TransformerEncoderLayer', agg_time_module='torch.nn.Identity', add_global_repr=False)
