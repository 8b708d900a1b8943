"""State-dict schema of `model.sync_model.Synchformer` under configs/sync.yaml (513 tensors, no buffers).

Mirrors SURVEY.md Appendix B; the names are the drop-in contract for `load_state_dict(ckpt['model'])`
(reference: model/sync_model.py:101-114, utils/logger.py:146, example.py:134).
"""
from collections import OrderedDict
from typing import Dict, Tuple

D = 768
V_DEPTH = 12
A_DEPTH = 12
S_DEPTH = 3
N_OFF_CLS = 21


def _ln(prefix: str, out: Dict[str, Tuple[int, ...]]):
    out[prefix + '.weight'] = (D,)
    out[prefix + '.bias'] = (D,)


def _lin(prefix: str, n_out: int, n_in: int, out: Dict[str, Tuple[int, ...]]):
    out[prefix + '.weight'] = (n_out, n_in)
    out[prefix + '.bias'] = (n_out,)


def _agg(prefix: str, out: Dict[str, Tuple[int, ...]]):
    # nn.TransformerEncoderLayer parameter order, then the cls_token registered by BaseEncoderLayer
    out[prefix + '.cls_token'] = (1, 1, D)
    out[prefix + '.self_attn.in_proj_weight'] = (3 * D, D)
    out[prefix + '.self_attn.in_proj_bias'] = (3 * D,)
    _lin(prefix + '.self_attn.out_proj', D, D, out)
    _lin(prefix + '.linear1', 4 * D, D, out)
    _lin(prefix + '.linear2', D, 4 * D, out)
    _ln(prefix + '.norm1', out)
    _ln(prefix + '.norm2', out)


def state_dict_schema(n_segments: int = 14, n_classes: int = N_OFF_CLS, head: str = 'off_head') -> 'OrderedDict[str, Tuple[int, ...]]':
    """name -> shape, for a sync transformer sequence of 2 + 14 * n_segments tokens."""
    s: Dict[str, Tuple[int, ...]] = OrderedDict()
    v = 'vfeat_extractor'
    s[f'{v}.cls_token'] = (1, 1, D)
    s[f'{v}.pos_embed'] = (1, 197, D)
    s[f'{v}.temp_embed'] = (1, 8, D)
    s[f'{v}.patch_embed.proj.weight'] = (D, 3, 16, 16)          # kept for strict loading; unused in forward
    s[f'{v}.patch_embed.proj.bias'] = (D,)
    s[f'{v}.patch_embed_3d.proj.weight'] = (D, 3, 2, 16, 16)
    s[f'{v}.patch_embed_3d.proj.bias'] = (D,)
    for i in range(V_DEPTH):
        b = f'{v}.blocks.{i}'
        _ln(f'{b}.norm1', s)
        _lin(f'{b}.attn.qkv', 3 * D, D, s)
        _lin(f'{b}.attn.proj', D, D, s)
        _lin(f'{b}.timeattn.qkv', 3 * D, D, s)
        _lin(f'{b}.timeattn.proj', D, D, s)
        _ln(f'{b}.norm2', s)
        _lin(f'{b}.mlp.fc1', 4 * D, D, s)
        _lin(f'{b}.mlp.fc2', D, 4 * D, s)
        _ln(f'{b}.norm3', s)
    _ln(f'{v}.norm', s)
    _agg(f'{v}.spatial_attn_agg', s)
    a = 'afeat_extractor'
    e = f'{a}.ast.embeddings'
    s[f'{e}.cls_token'] = (1, 1, D)
    s[f'{e}.distillation_token'] = (1, 1, D)
    s[f'{e}.position_embeddings'] = (1, 74, D)
    s[f'{e}.patch_embeddings.projection.weight'] = (D, 1, 16, 16)
    s[f'{e}.patch_embeddings.projection.bias'] = (D,)
    for i in range(A_DEPTH):
        l = f'{a}.ast.encoder.layer.{i}'
        for n in ('query', 'key', 'value'):
            _lin(f'{l}.attention.attention.{n}', D, D, s)
        _lin(f'{l}.attention.output.dense', D, D, s)
        _lin(f'{l}.intermediate.dense', 4 * D, D, s)
        _lin(f'{l}.output.dense', D, 4 * D, s)
        _ln(f'{l}.layernorm_before', s)
        _ln(f'{l}.layernorm_after', s)
    _ln(f'{a}.ast.layernorm', s)
    _agg(f'{a}.freq_attn_agg', s)
    _lin('vproj', D, D, s)
    _lin('aproj', D, D, s)
    t = 'transformer'
    s[f'{t}.OFF_tok'] = (1, 1, D)
    s[f'{t}.MOD_tok'] = (1, 1, D)
    _ln(f'{t}.vis_in_lnorm', s)
    _ln(f'{t}.aud_in_lnorm', s)
    s[f'{t}.pos_emb_cfg.pos_emb'] = (1, 2 + 14 * n_segments, D)
    for i in range(S_DEPTH):
        b = f'{t}.blocks.{i}'
        _ln(f'{b}.ln1', s)
        _ln(f'{b}.ln2', s)
        for n in ('key', 'query', 'value', 'proj'):
            _lin(f'{b}.attn.{n}', D, D, s)
        _lin(f'{b}.mlp.0', 4 * D, D, s)
        _lin(f'{b}.mlp.2', D, 4 * D, s)
    _ln(f'{t}.ln_f', s)
    _lin(f'{t}.{head}', n_classes, D, s)
    return s
