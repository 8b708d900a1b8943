"""Import the UNMODIFIED reference (read-only at /root/reference) inside THIS container.

Test infrastructure only: used by `make_golden.py` to produce the committed fixtures and by
the optional `-m "not gpu"` cross-check `tests/test_oracle_vs_reference.py` (skipped when
/root/reference is absent, e.g. on the GPU box).  Nothing in the product imports this.

The reference needs three third-party packages that are not in the image:
  * omegaconf  -> a tiny attribute-dict stub with `OmegaConf.load/create` and `${a.b}` resolve
  * timm       -> only `trunc_normal_`, `DropPath`, `to_2tuple` are *used*
  * transformers 4.27 APIs removed in 5.x -> `find_pruneable_heads_and_indices`, `get_head_mask`
The stubs are created as in-memory modules; no file of the reference is copied or modified.
"""
import importlib
import os
import re
import sys
import types

REF_ROOT = os.environ.get('SYNCHFORMER_REF', '/root/reference')


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, 'model'))


class _AttrDict(dict):
    """dict with attribute access (enough of DictConfig for the model constructors)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def __delattr__(self, k):
        del self[k]


def _wrap(o):
    if isinstance(o, dict):
        return _AttrDict({k: _wrap(v) for k, v in o.items()})
    if isinstance(o, (list, tuple)):
        return [_wrap(v) for v in o]
    return o


def _resolve(root, node=None):
    """Resolve `${a.b.c}` interpolations in place (whole-string references only)."""
    node = root if node is None else node
    items = node.items() if isinstance(node, dict) else enumerate(node)
    for k, v in list(items):
        if isinstance(v, (dict, list)):
            _resolve(root, v)
        elif isinstance(v, str):
            m = re.fullmatch(r'\$\{([A-Za-z0-9_.]+)\}', v)
            if m:
                cur = root
                try:
                    for part in m.group(1).split('.'):
                        cur = cur[part]
                except (KeyError, TypeError):
                    continue
                node[k] = cur
    return root


def _install_omegaconf_stub():
    if 'omegaconf' in sys.modules:
        return
    import yaml
    mod = types.ModuleType('omegaconf')

    class OmegaConf:
        @staticmethod
        def load(path):
            with open(path) as f:
                return _resolve(_wrap(yaml.safe_load(f)))

        @staticmethod
        def create(obj=None):
            return _resolve(_wrap(obj or {}))

        @staticmethod
        def to_container(cfg, resolve=True):
            return cfg

        @staticmethod
        def resolve(cfg):
            _resolve(cfg)

    mod.OmegaConf = OmegaConf
    mod.DictConfig = _AttrDict
    mod.ListConfig = list
    mod.dictconfig = types.ModuleType('omegaconf.dictconfig')
    mod.dictconfig.DictConfig = _AttrDict
    sys.modules['omegaconf'] = mod
    sys.modules['omegaconf.dictconfig'] = mod.dictconfig


def _install_timm_stub():
    if 'timm' in sys.modules:
        return
    import torch
    timm = types.ModuleType('timm')
    models = types.ModuleType('timm.models')
    layers = types.ModuleType('timm.models.layers')
    resnet = types.ModuleType('timm.models.resnet')
    registry = types.ModuleType('timm.models.registry')
    data = types.ModuleType('timm.data')

    class DropPath(torch.nn.Module):  # identity in eval, which is all the oracle runs
        def __init__(self, drop_prob=0.0):
            super().__init__()
            self.drop_prob = drop_prob

        def forward(self, x):
            assert not self.training or self.drop_prob == 0.0, 'stub DropPath: eval only'
            return x

    layers.trunc_normal_ = torch.nn.init.trunc_normal_
    layers.DropPath = DropPath
    layers.to_2tuple = lambda x: x if isinstance(x, tuple) else (x, x)
    resnet.resnet26d = resnet.resnet50d = lambda *a, **k: None
    registry.register_model = lambda fn: fn
    data.IMAGENET_DEFAULT_MEAN = (0.485, 0.456, 0.406)
    data.IMAGENET_DEFAULT_STD = (0.229, 0.224, 0.225)
    timm.models, timm.data = models, data
    models.layers, models.resnet, models.registry = layers, resnet, registry
    from importlib.machinery import ModuleSpec
    for name, m in [('timm', timm), ('timm.models', models), ('timm.models.layers', layers),
                    ('timm.models.resnet', resnet), ('timm.models.registry', registry), ('timm.data', data)]:
        m.__spec__ = ModuleSpec(name, loader=None)  # transformers probes find_spec('timm')
        sys.modules[name] = m


def _patch_transformers():
    import transformers.pytorch_utils as pu
    if not hasattr(pu, 'find_pruneable_heads_and_indices'):
        pu.find_pruneable_heads_and_indices = lambda *a, **k: (set(), None)
    if not hasattr(pu, 'prune_linear_layer'):
        pu.prune_linear_layer = lambda layer, *a, **k: layer
    from transformers.modeling_utils import PreTrainedModel
    if not hasattr(PreTrainedModel, 'get_head_mask'):
        PreTrainedModel.get_head_mask = lambda self, head_mask, n, *a, **k: [None] * n


def import_reference():
    """Returns (Synchformer class, sync.yaml model-config loader).  CWD is switched to the reference root
    because the reference uses relative sys.path entries (model/sync_model.py:9, visual/__init__.py:2)."""
    assert reference_available(), f'reference not found at {REF_ROOT}'
    _patch_transformers()
    _install_omegaconf_stub()
    _install_timm_stub()
    os.chdir(REF_ROOT)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    sync_model = importlib.import_module('model.sync_model')
    from omegaconf import OmegaConf

    def load_model_cfg(n_segments: int = 14):
        cfg = OmegaConf.load(os.path.join(REF_ROOT, 'configs', 'sync.yaml'))
        cfg.model.params.transformer.params.pos_emb_cfg.params.block_shape = [2 + 14 * n_segments]
        cfg.model.params.transformer.params.off_head_cfg.params.out_features = 21
        return cfg.model

    return sync_model, load_model_cfg


def build_reference_model(n_segments: int = 14):
    import torch
    sync_model, load_model_cfg = import_reference()
    mcfg = load_model_cfg(n_segments)
    model = sync_model.Synchformer(**mcfg['params'])
    model.eval()
    for p in model.parameters():
        p.requires_grad_(False)
    return model
