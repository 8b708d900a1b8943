"""Host-side mirror of the reference's model classes for the hot path, driving the sm_100a kernels.

Same constructor kwargs, attribute names, `forward()` signature and `state_dict()` schema as
`model.sync_model.Synchformer` (reference model/sync_model.py:23-114) and the modules it instantiates:
  MotionFormer  <- model/modules/feat_extractors/visual/motionformer.py:24-272 (divided space-time ViT-B + spatial CLS aggregator)
  AST           <- model/modules/feat_extractors/audio/ast.py:13-250           (HF AST encoder + frequency CLS aggregator)
  GlobalTransformer[WithSyncabilityHead] <- model/sync_model.py:117-190
Parameters are ordinary fp32 `nn.Parameter`s under the reference's names (so `load_state_dict(ckpt['model'])`,
`.to(device)`, optimisers and DDP keep working); bf16 copies of the GEMM weights (q/k/v fused where the reference
keeps three matrices) are cached per device and refreshed when a parameter's version counter changes.

All arithmetic runs in the CUDA kernels of `libsynchformer_b200.so` (see ops.py); PyTorch only owns memory and
streams.  Everything also trains: in train mode the synchronisation module (vproj / aproj / transformer) applies dropout and is
differentiable (train.py), and so are the two feature extractors when they have trainable parameters
(train_encoders.py, §8f N1); frozen / eval towers, as configs/sync.yaml uses them, keep the inference path.
"""
import logging
import math
import os
from typing import Any, Dict, Mapping, Optional

import torch
from torch import nn

from . import ops, train, train_encoders
from .schema import D, state_dict_schema

EPS_V, EPS_A, EPS_S = 1e-6, 1e-12, 1e-5
V_TOK, V_SPACE, V_FRAMES = 1569, 196, 8
A_TOK, A_F, A_T = 74, 12, 6


class _Params(nn.Module):
    """Plain parameter container (a node of the reference's module tree)."""


def _build_tree(root: nn.Module, schema: Mapping[str, tuple]):
    for name, shape in schema.items():
        parts = name.split('.')
        mod = root
        for p in parts[:-1]:
            if p not in mod._modules:
                mod.add_module(p, _Params())
            mod = mod._modules[p]
        mod.register_parameter(parts[-1], nn.Parameter(torch.zeros(shape)))


def _init_reference_like(module: nn.Module):
    """Random init in the spirit of the reference (trunc_normal(.02) linears, unit LayerNorm, small tokens; the Conv3d
    patch embedding stays zero exactly as video_model_builder.py:61 leaves it)."""
    with torch.no_grad():
        for name, p in module.named_parameters():
            leaf = name.rsplit('.', 1)[-1]
            is_ln = any(t in name for t in ('norm', 'lnorm', 'ln1', 'ln2', 'ln_f'))
            if is_ln:
                p.fill_(1.0 if leaf == 'weight' else 0.0)
            elif leaf in ('bias', 'in_proj_bias'):
                p.zero_()
            elif 'patch_embed_3d.proj.weight' in name:
                p.zero_()
            elif leaf in ('OFF_tok', 'MOD_tok', 'pos_emb'):
                p.normal_(0.0, 1.0)                       # torch.randn in sync_model.py:129-130, transformer.py:126
            else:
                nn.init.trunc_normal_(p, std=0.02)


def _init_from_stage1_ckpt(module: nn.Module, ckpt_path: str, encoder: str, what: str):
    """Stage-I (AVCLIP) checkpoint -> feature-extractor initialisation, as motionformer.py:156-173 / ast.py:113-132 do when `ckpt_path`
    ends with '.pt' (scripts/sbatch_train_sync.sh:74-75 launches stage II this way): keep the `[module.]{v,a}_encoder.` entries of
    `ckpt['state_dict']`, strip the prefix, `load_state_dict(strict=False)`, and warn about missing / unexpected keys.
    Other values of ckpt_path (the HF hub name of the AudioSet AST, the SSv2 `.pyth` Motionformer files) are public pre-trained
    initialisations the reference downloads; there is no network here, so they are refused - load such weights with load_state_dict."""
    if not str(ckpt_path).endswith('.pt'):
        raise NotImplementedError(f"{what}: ckpt_path={ckpt_path!r} is a downloadable pre-trained initialisation (harness work, needs the network); "
                                  "only stage-I '.pt' checkpoints are initialised from here - load other weights with load_state_dict")
    if not os.path.exists(ckpt_path):
        raise ValueError(f'Cant find the checkpoint file: {ckpt_path}.', 'Please download it manually and ensure the path exists.')   # utils/utils.py:57-58
    try:
        ckpt = torch.load(ckpt_path, map_location='cpu', weights_only=True)
    except Exception:                                   # reference checkpoints also pickle their OmegaConf `args`
        ckpt = torch.load(ckpt_path, map_location='cpu', weights_only=False)
    weights = {}
    for k, v in ckpt['state_dict'].items():
        if k.startswith((f'module.{encoder}.', f'{encoder}.')):
            weights[k.replace('module.', '').replace(f'{encoder}.', '')] = v
    status = module.load_state_dict(weights, strict=False)
    if len(status.missing_keys) > 0 or len(status.unexpected_keys) > 0:
        logging.warning(f'Loading exact {what} ckpt from {ckpt_path} failed. \nMissing keys ({len(status.missing_keys)}): {status.missing_keys}, \n'
                        f'Unexpected keys ({len(status.unexpected_keys)}): {status.unexpected_keys} \n'
                        'temp_attn_agg are expected to be missing if ckpt was pt contrastively.')
    else:
        logging.info(f'Loading {what} ckpt from {ckpt_path} succeeded.')
    return status


class _MeanTokens(torch.autograd.Function):
    """AveragePooling 'BS T D -> BS D' (motionformer.py:405-409) over the 8 / 6 time tokens of a segment."""

    @staticmethod
    def forward(ctx, x):
        ctx.T = x.shape[-2]
        return ops.mean_tokens(x.float().contiguous().view(-1, x.shape[-2], x.shape[-1])).view(*x.shape[:-2], x.shape[-1])

    @staticmethod
    def backward(ctx, g):
        return ops.mean_tokens_bwd(g.float().contiguous().view(-1, g.shape[-1]), ctx.T).view(*g.shape[:-1], ctx.T, g.shape[-1])


def _tower_trains(m: nn.Module) -> bool:
    """A feature extractor takes the differentiable path iff it is in train mode, autograd is on and it has trainable parameters
    (stage I, or stage II with `is_trainable: True`); frozen / eval towers keep the inference path."""
    return m.training and torch.is_grad_enabled() and any(p.requires_grad for n, p in m.named_parameters() if not n.startswith('patch_embed.'))


class _KernelModule(nn.Module):
    """Shared machinery: bf16 weight cache keyed on (device, parameter versions)."""

    def __init__(self):
        super().__init__()
        self._wcache: Dict[str, torch.Tensor] = {}
        self._wcache_key = None
        self._taps: Optional[Dict[str, torch.Tensor]] = None     # parity tests set a dict here to receive copies of the residual stream

    def _tap(self, name: str, x: torch.Tensor, rows_per_item: int):
        """Diagnostic tap (tests only; None on the product path): fp32 copy of `x` viewed as (items, rows_per_item, 768), under the names
        tests/golden/make_golden.py and the oracle use."""
        if self._taps is not None:
            t = x.detach().float().clone().view(-1, rows_per_item, D)
            self._taps[name] = torch.cat([self._taps[name], t], dim=0) if name in self._taps else t

    def _own_params(self) -> Dict[str, nn.Parameter]:
        return dict(self.named_parameters())

    def _pack(self, P: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        raise NotImplementedError

    def weights(self):
        """(fp32 parameter dict, bf16 GEMM-weight dict), both on the module's CUDA device."""
        P = self._own_params()
        first = next(iter(P.values()))
        ops.require_cuda(first, type(self).__name__ + ' parameters')
        key = (first.device, tuple((p._version, p.data_ptr()) for p in P.values()))
        if key != self._wcache_key:
            for n, p in P.items():
                if p.dtype != torch.float32 or not p.is_contiguous():
                    raise TypeError(f'{n}: parameters must be contiguous fp32 (got {p.dtype})')
            with torch.no_grad():
                self._wcache = self._pack({n: p.detach() for n, p in P.items()})
            self._wcache_key = key
        return {n: p.detach() for n, p in P.items()}, self._wcache

    @staticmethod
    def _bf16(w: torch.Tensor) -> torch.Tensor:
        return ops.cast_bf16(w.reshape(w.shape[0], -1).contiguous())


def _cls_aggregator(P, W, prefix: str, kv_src: torch.Tensor, n_groups_outer: int, n_inner: int, kv_outer_rows: int, kv_inner_rows: int,
                    kv_row_rows: int, Lk: int) -> torch.Tensor:
    """CLS-row-only evaluation of BaseEncoderLayer (motionformer.py:301-334): the layer returns x[:, 0] only, so queries,
    out-proj and the FFN are needed for the CLS row alone; keys/values need every token.  Exact (no approximation).
    kv_src: (rows, 768) bf16 = norm1(final_norm(tokens)); returns (n_groups_outer * n_inner, 768) fp32."""
    dev = kv_src.device
    G = n_groups_outer * n_inner
    kv = ops.gemm(kv_src, W[prefix + 'kv_w'], P[prefix + 'self_attn.in_proj_bias'][D:])                                   # (rows, 1536) bf16
    # the CLS token is a learned constant -> its norm1 / q / k / v rows are computed once per call (M = 1 GEMM)
    cls_ln = ops.layernorm(P[prefix + 'cls_token'].view(1, D), P[prefix + 'norm1.weight'], P[prefix + 'norm1.bias'], EPS_V)
    cls_qkv = ops.gemm(cls_ln, W[prefix + 'in_w'], P[prefix + 'self_attn.in_proj_bias'])              # (1, 2304) bf16
    ao = ops.empty_bf16((G, D), dev)
    ops.attention(cls_qkv, kv, kv[:, D:], ao, q_strides=(0, 0, 0),
                  kv_strides=(kv_outer_rows * 2 * D, kv_inner_rows * 2 * D, kv_row_rows * 2 * D),
                  o_strides=(n_inner * D, D, D), n_outer=n_groups_outer, n_inner=n_inner, n_heads=12, head_dim=64, Lq=1, Lk=Lk,
                  scale=0.125, k_prefix=cls_qkv[:, D:], v_prefix=cls_qkv[:, 2 * D:], prefix_outer=0)
    y0 = ops.gemm(ao, W[prefix + 'out_w'], P[prefix + 'self_attn.out_proj.bias'], residual=P[prefix + 'cls_token'].view(1, D), out_f32=True)
    h = ops.layernorm(y0, P[prefix + 'norm2.weight'], P[prefix + 'norm2.bias'], EPS_V)
    h = ops.gemm(h, W[prefix + 'l1_w'], P[prefix + 'linear1.bias'], gelu=True)
    return ops.gemm(h, W[prefix + 'l2_w'], P[prefix + 'linear2.bias'], residual=y0, out_f32=True)


def _pack_aggregator(P, prefix: str, out: Dict[str, torch.Tensor], bf16):
    in_w = P[prefix + 'self_attn.in_proj_weight']
    out[prefix + 'in_w'] = bf16(in_w)
    out[prefix + 'kv_w'] = bf16(in_w[D:])
    out[prefix + 'out_w'] = bf16(P[prefix + 'self_attn.out_proj.weight'])
    out[prefix + 'l1_w'] = bf16(P[prefix + 'linear1.weight'])
    out[prefix + 'l2_w'] = bf16(P[prefix + 'linear2.weight'])


class MotionFormer(_KernelModule):
    """Visual stream.  Constructor mirrors motionformer.py:39-47; only the configuration the sync / AVCLIP configs use is
    implemented: extract_features=True, factorize_space_time=True, agg_space_module='TransformerEncoderLayer',
    agg_time_module in {Identity, 'AveragePooling'}, add_global_repr=False, divided space-time attention."""

    def __init__(self, extract_features: bool = False, ckpt_path: str = None, factorize_space_time: bool = None,
                 agg_space_module: str = None, agg_time_module: str = None, add_global_repr: bool = True,
                 agg_segments_module: str = None, max_segments: int = None):
        super().__init__()
        if not extract_features or not factorize_space_time or agg_space_module != 'TransformerEncoderLayer' or add_global_repr:
            raise NotImplementedError('synchformer_b200.MotionFormer supports extract_features=True, factorize_space_time=True, '
                                      "agg_space_module='TransformerEncoderLayer', add_global_repr=False (configs/sync.yaml:18-27)")
        self.ckpt_path = ckpt_path
        self.time_pool = 'AveragePooling' in str(agg_time_module)
        if not self.time_pool and 'Identity' not in str(agg_time_module):
            raise NotImplementedError(f'agg_time_module={agg_time_module}')
        self.extract_features, self.factorize_space_time, self.add_global_repr = True, True, False
        self.embed_dim, self.num_heads = D, 12
        self.max_segments_per_pass = 512
        # LayerNorm fused into the GEMMs on either side of it (csrc/gemm_tcgen05.cu: EMIT_LN / LN_FOLD) instead of 36 LayerNorm launches per
        # pass.  Opt-in (SFB_LN_FUSED=1 / 2): measured on the B200 the fused schedules remove 13 - 21 ms of LayerNorm passes per 64-clip step and
        # add as much to the GEMM epilogues (profiles/r2_ab_f32_tma_epilogue.txt, r2_ab_ln_fusion_modes.txt).
        # 2 = only the two norms in front of the qkv GEMMs (norm3, norm1) are fused; norm2 keeps its LayerNorm launch, because the folded
        # epilogue costs 0.16 ms on a qkv GEMM but 0.66 ms on fc1, whose epilogue already carries the GELU (tools/gemm_ab.py)
        self.fuse_layernorm = int(os.environ.get('SFB_LN_FUSED', '0'))
        schema = {k[len('vfeat_extractor.'):]: v for k, v in state_dict_schema().items() if k.startswith('vfeat_extractor.')}
        _build_tree(self, schema)
        _init_reference_like(self)
        if ckpt_path is not None:
            _init_from_stage1_ckpt(self, ckpt_path, 'v_encoder', 'vfeat_extractor')
        self.patch_embed.requires_grad_(False)                       # motionformer.py:177

    def _pack(self, P):
        W = {'pe_w': self._bf16(P['patch_embed_3d.proj.weight'])}
        for i in range(12):
            b = f'blocks.{i}.'
            for n in ('attn.qkv', 'attn.proj', 'timeattn.qkv', 'timeattn.proj', 'mlp.fc1', 'mlp.fc2'):
                W[b + n] = self._bf16(P[b + n + '.weight'])
            # LayerNorm folded into the Linear that consumes it (vit_helper.py:366-375: norm3 -> timeattn.qkv, norm1 -> attn.qkv, norm2 -> mlp.fc1):
            #   LN(x) W^T + b = rstd (x (gamma . W)^T - mean colsum) + (b + W beta)
            for norm, lin in (('norm3', 'timeattn.qkv'), ('norm1', 'attn.qkv'), ('norm2', 'mlp.fc1')):
                w0, b0 = P[b + lin + '.weight'], P[b + lin + '.bias']
                gamma, beta = P[b + norm + '.weight'], P[b + norm + '.bias']
                wf = self._bf16(w0 * gamma.unsqueeze(0))
                W[b + lin + '.fold_w'] = wf
                W[b + lin + '.fold_cs'] = wf.float().sum(dim=1).contiguous()          # sums of the weights AS THE TENSOR CORES SEE THEM
                W[b + lin + '.fold_b'] = (b0 + w0 @ beta).contiguous()
        _pack_aggregator(P, 'spatial_attn_agg.', W, self._bf16)
        return W

    def _divided_attention(self, qkv: torch.Tensor, att: torch.Tensor, n: int, mode: str):
        """DividedAttention.forward vit_helper.py:100-158 on the fused (n*1569, 2304) qkv activations."""
        row, seg = 3 * D, V_TOK * 3 * D
        q, k, v = qkv, qkv[:, D:], qkv[:, 2 * D:]
        q1, k1, v1, o1 = qkv[1:], qkv[1:, D:], qkv[1:, 2 * D:], att[1:]
        if mode == 'time':    # '(b n) f d': 8 frames of one location + CLS key/value
            fused = ops.attention(q1, k1, v1, o1, q_strides=(seg, row, V_SPACE * row), kv_strides=(seg, row, V_SPACE * row),
                                  o_strides=(V_TOK * D, D, V_SPACE * D), n_outer=n, n_inner=V_SPACE, n_heads=12, head_dim=64, Lq=V_FRAMES,
                                  Lk=V_FRAMES, scale=0.125, k_prefix=k, v_prefix=v, prefix_outer=seg,
                                  q_extra=q, q_extra_outer=seg, extra_out=att, extra_out_outer=V_TOK * D)    # fused only with SFB_TIME_CLS_FUSED=1
        else:                 # '(b f) n d': 196 locations of one frame + CLS key/value; the CLS query rides along (fused) when supported
            fused = ops.attention(q1, k1, v1, o1, q_strides=(seg, V_SPACE * row, row), kv_strides=(seg, V_SPACE * row, row),
                                  o_strides=(V_TOK * D, V_SPACE * D, D), n_outer=n, n_inner=V_FRAMES, n_heads=12, head_dim=64, Lq=V_SPACE,
                                  Lk=V_SPACE, scale=0.125, k_prefix=k, v_prefix=v, prefix_outer=seg,
                                  q_extra=q, q_extra_outer=seg, extra_out=att, extra_out_outer=V_TOK * D)
        if not fused:         # CLS query attends to all 1569 keys (:124)
            ops.attention(q, k, v, att, q_strides=(seg, 0, row), kv_strides=(seg, 0, row), o_strides=(V_TOK * D, 0, D), n_outer=n, n_inner=1,
                          n_heads=12, head_dim=64, Lq=1, Lk=V_TOK, scale=0.125)

    def _encode_chunk(self, vis: Optional[torch.Tensor], P, W, a: Optional[torch.Tensor] = None) -> torch.Tensor:
        """vis (n, 16, 3, 224, 224) [or its im2col matrix `a` (n*1568, 1536)] -> (n, 8, 768) fp32."""
        if a is None:
            a = ops.im2col_video(vis)
        n = a.shape[0] // 1568
        dev = a.device
        patch = ops.gemm(a, W['pe_w'], P['patch_embed_3d.proj.bias'], out_f32=True)
        x = ops.video_tokens(patch, P['cls_token'], P['pos_embed'], P['temp_embed'], n)           # (n*1569, 768) fp32 residual stream
        del a, patch
        self._tap('v_embed', x, V_TOK)
        M = n * V_TOK
        ln, qkv, att, hid = (ops.empty_bf16((M, w * D), dev) for w in (1, 3, 1, 4))
        fused = self.fuse_layernorm if ops.GEMM_IMPL != 1 else 0    # the CUDA-core bring-up GEMM has no EMIT_LN epilogue
        if fused == 2:
            # norm3 / norm1 fused into the qkv GEMMs (statistics and bf16 copy from the fc2 / time-proj epilogues), norm2 as a LayerNorm launch
            st0 = torch.empty((M, 1, 2), device=dev, dtype=torch.float32)
            st = torch.empty((M, D // 64, 2), device=dev, dtype=torch.float32)
            ops.rowstats_cast(x, ln, st0)
            cur = st0
            for i in range(12):
                b = f'blocks.{i}.'
                ops.gemm(ln, W[b + 'timeattn.qkv.fold_w'], W[b + 'timeattn.qkv.fold_b'], out=qkv, ln_fold=(cur, W[b + 'timeattn.qkv.fold_cs'], EPS_V))
                self._divided_attention(qkv, att, n, 'time')
                ops.gemm(att, W[b + 'timeattn.proj'], P[b + 'timeattn.proj.bias'], out=x, residual=x, out_f32=True, emit_ln=(ln, st))
                cur = st
                ops.gemm(ln, W[b + 'attn.qkv.fold_w'], W[b + 'attn.qkv.fold_b'], out=qkv, ln_fold=(st, W[b + 'attn.qkv.fold_cs'], EPS_V))
                self._divided_attention(qkv, att, n, 'space')
                ops.gemm(att, W[b + 'attn.proj'], P[b + 'attn.proj.bias'], out=x, residual=x, out_f32=True)
                ops.layernorm(x, P[b + 'norm2.weight'], P[b + 'norm2.bias'], EPS_V, out=ln)
                ops.gemm(ln, W[b + 'mlp.fc1'], P[b + 'mlp.fc1.bias'], out=hid, gelu=True)
                ops.gemm(hid, W[b + 'mlp.fc2'], P[b + 'mlp.fc2.bias'], out=x, residual=x, out_f32=True, emit_ln=(ln, st) if i < 11 else None)
                if i in (0, 11):
                    self._tap(f'v_block{i}', x, V_TOK)
        elif fused:
            # No LayerNorm pass inside the blocks: every residual GEMM (proj / fc2) leaves a bf16 copy of the stream in `ln` plus per-row
            # partial sums in `st`, and the next qkv / fc1 GEMM normalises in its epilogue (folded gamma / beta).
            st0 = torch.empty((M, 1, 2), device=dev, dtype=torch.float32)
            st = torch.empty((M, D // 64, 2), device=dev, dtype=torch.float32)
            ops.rowstats_cast(x, ln, st0)
            cur = st0
            for i in range(12):                                                                # DividedSpaceTimeBlock vit_helper.py:364-376
                b = f'blocks.{i}.'
                ops.gemm(ln, W[b + 'timeattn.qkv.fold_w'], W[b + 'timeattn.qkv.fold_b'], out=qkv, ln_fold=(cur, W[b + 'timeattn.qkv.fold_cs'], EPS_V))
                self._divided_attention(qkv, att, n, 'time')
                ops.gemm(att, W[b + 'timeattn.proj'], P[b + 'timeattn.proj.bias'], out=x, residual=x, out_f32=True, emit_ln=(ln, st))
                cur = st
                ops.gemm(ln, W[b + 'attn.qkv.fold_w'], W[b + 'attn.qkv.fold_b'], out=qkv, ln_fold=(st, W[b + 'attn.qkv.fold_cs'], EPS_V))
                self._divided_attention(qkv, att, n, 'space')
                ops.gemm(att, W[b + 'attn.proj'], P[b + 'attn.proj.bias'], out=x, residual=x, out_f32=True, emit_ln=(ln, st))
                ops.gemm(ln, W[b + 'mlp.fc1.fold_w'], W[b + 'mlp.fc1.fold_b'], out=hid, gelu=True, ln_fold=(st, W[b + 'mlp.fc1.fold_cs'], EPS_V))
                ops.gemm(hid, W[b + 'mlp.fc2'], P[b + 'mlp.fc2.bias'], out=x, residual=x, out_f32=True, emit_ln=(ln, st) if i < 11 else None)
                if i in (0, 11):
                    self._tap(f'v_block{i}', x, V_TOK)
        for i in range(0 if fused else 12):                                                    # one LayerNorm launch per norm (default)
            b = f'blocks.{i}.'
            ops.layernorm(x, P[b + 'norm3.weight'], P[b + 'norm3.bias'], EPS_V, out=ln)
            ops.gemm(ln, W[b + 'timeattn.qkv'], P[b + 'timeattn.qkv.bias'], out=qkv)
            self._divided_attention(qkv, att, n, 'time')
            ops.gemm(att, W[b + 'timeattn.proj'], P[b + 'timeattn.proj.bias'], out=x, residual=x, out_f32=True)
            ops.layernorm(x, P[b + 'norm1.weight'], P[b + 'norm1.bias'], EPS_V, out=ln)
            ops.gemm(ln, W[b + 'attn.qkv'], P[b + 'attn.qkv.bias'], out=qkv)
            self._divided_attention(qkv, att, n, 'space')
            ops.gemm(att, W[b + 'attn.proj'], P[b + 'attn.proj.bias'], out=x, residual=x, out_f32=True)
            ops.layernorm(x, P[b + 'norm2.weight'], P[b + 'norm2.bias'], EPS_V, out=ln)
            ops.gemm(ln, W[b + 'mlp.fc1'], P[b + 'mlp.fc1.bias'], out=hid, gelu=True)
            ops.gemm(hid, W[b + 'mlp.fc2'], P[b + 'mlp.fc2.bias'], out=x, residual=x, out_f32=True)
            if i in (0, 11):
                self._tap(f'v_block{i}', x, V_TOK)
        # final norm on the 1568 non-CLS tokens (motionformer.py:229-232) fused with the aggregator's norm1
        g = 'spatial_attn_agg.'
        kv_src = ops.layernorm(x, P['norm.weight'], P['norm.bias'], EPS_V, out=ln[:n * 1568], rows=n * 1568, group=1568, group_stride=V_TOK,
                               offset=1, gamma2=P[g + 'norm1.weight'], beta2=P[g + 'norm1.bias'], eps2=EPS_V)
        feats = _cls_aggregator(P, W, g, kv_src, n, V_FRAMES, 1568, V_SPACE, 1, V_SPACE)          # (n*8, 768)
        return feats.view(n, V_FRAMES, D)

    def forward(self, x: torch.Tensor, for_loop: bool = False, cont_mask: torch.Tensor = None):
        """x (B, S, C=3, T=16, H, W) as in motionformer.py:182-223 (a permuted view of the (B,S,T,C,H,W) input).
        Returns ((B, S, 8, 768) or (B, S, 768) with AveragePooling, None)."""
        if cont_mask is not None:
            raise NotImplementedError('cont_mask is not supported (no caller in the reference passes it)')
        return self.encode(x.permute(0, 1, 3, 2, 4, 5)), None

    def encode_clip(self, clip: torch.Tensor, n_segments: int, v_start: int, v_stride: int) -> torch.Tensor:
        """N2: clip (B, n_frames, 3, 224, 224), any supported dtype incl. raw uint8 -> (B, S, 8, 768); the S overlapping 16-frame
        segments (GenerateMultipleSegments, dataset/transforms.py:402-499) are sliced inside the patch-embedding gather."""
        ops.require_cuda(clip, 'clip')
        B = clip.shape[0]
        P, W = self.weights()
        per = max(1, self.max_segments_per_pass // n_segments)               # whole clips per pass
        outs = [self._encode_chunk(None, P, W, a=ops.im2col_video_clip(clip[b:b + per].contiguous(), n_segments, v_start, v_stride))
                for b in range(0, B, per)]
        feats = (outs[0] if len(outs) == 1 else torch.cat(outs, dim=0)).view(B, n_segments, V_FRAMES, D)
        return _MeanTokens.apply(feats) if self.time_pool else feats

    def encode(self, vis: torch.Tensor) -> torch.Tensor:
        """vis (B, S, T=16, C=3, 224, 224) fp32 / fp16 / bf16 / uint8 -> (B, S, 8, 768) fp32."""
        ops.require_cuda(vis, 'vis')
        if vis.dim() != 6 or tuple(vis.shape[2:]) != (16, 3, 224, 224):
            raise ValueError(f'expected video of shape (B, S, 16, 3, 224, 224), got {tuple(vis.shape)}')
        B, S = vis.shape[:2]
        if _tower_trains(self):                                  # differentiable forward + hand-written backward
            return train_encoders.motionformer_features(self, vis)
        P, W = self.weights()
        flat = vis.contiguous().view(B * S, 16, 3, 224, 224)
        outs = [self._encode_chunk(flat[s:s + self.max_segments_per_pass], P, W) for s in range(0, B * S, self.max_segments_per_pass)]
        feats = outs[0] if len(outs) == 1 else torch.cat(outs, dim=0)
        feats = feats.view(B, S, V_FRAMES, D)
        return _MeanTokens.apply(feats) if self.time_pool else feats


class AST(_KernelModule):
    """Audio stream.  Constructor mirrors ast.py:14-27; supported: extract_features=True, factorize_freq_time=True,
    agg_freq_module='TransformerEncoderLayer', agg_time_module in {Identity, 'AveragePooling'}, add_global_repr=False, max_spec_t=66."""

    def __init__(self, extract_features: bool = False, ckpt_path: str = None, feat_type: str = None, max_spec_t: int = None,
                 factorize_freq_time: bool = None, agg_freq_module: str = None, agg_time_module: str = None, add_global_repr: bool = True,
                 agg_segments_module: str = None, max_segments: int = None):
        super().__init__()
        if not extract_features or not factorize_freq_time or agg_freq_module != 'TransformerEncoderLayer' or add_global_repr:
            raise NotImplementedError('synchformer_b200.AST supports extract_features=True, factorize_freq_time=True, '
                                      "agg_freq_module='TransformerEncoderLayer', add_global_repr=False (configs/sync.yaml:6-17)")
        self.ckpt_path = ckpt_path
        if max_spec_t not in (None, 66):
            raise NotImplementedError('max_spec_t must be 66 (74 position embeddings)')
        self.time_pool = 'AveragePooling' in str(agg_time_module)
        if not self.time_pool and 'Identity' not in str(agg_time_module):
            raise NotImplementedError(f'agg_time_module={agg_time_module}')
        self.extract_features, self.factorize_freq_time, self.add_global_repr, self.max_spec_t = True, True, False, 66
        schema = {k[len('afeat_extractor.'):]: v for k, v in state_dict_schema().items() if k.startswith('afeat_extractor.')}
        _build_tree(self, schema)
        _init_reference_like(self)
        if ckpt_path is not None:
            _init_from_stage1_ckpt(self, ckpt_path, 'a_encoder', 'afeat_extractor')

    def _pack(self, P):
        W = {'pe_w': self._bf16(P['ast.embeddings.patch_embeddings.projection.weight'])}
        self._fused_bias = {}
        for i in range(12):
            l = f'ast.encoder.layer.{i}.'
            a = l + 'attention.attention.'
            W[l + 'qkv'] = self._bf16(torch.cat([P[a + 'query.weight'], P[a + 'key.weight'], P[a + 'value.weight']], dim=0))
            W[l + 'qkv_b'] = torch.cat([P[a + 'query.bias'], P[a + 'key.bias'], P[a + 'value.bias']], dim=0).contiguous()
            W[l + 'o'] = self._bf16(P[l + 'attention.output.dense.weight'])
            W[l + 'fc1'] = self._bf16(P[l + 'intermediate.dense.weight'])
            W[l + 'fc2'] = self._bf16(P[l + 'output.dense.weight'])
        _pack_aggregator(P, 'freq_attn_agg.', W, self._bf16)
        return W

    def forward(self, x: torch.Tensor, for_loop: bool = False, cont_mask: torch.Tensor = None, **ast_kwargs):
        """x (B, S, T=66, F=128) as in ast.py:137-176 (a permuted view of (B, S, F, T)).  Returns ((B,S,6,768) | (B,S,768), None)."""
        if cont_mask is not None:
            raise NotImplementedError('cont_mask is not supported (no caller in the reference passes it)')
        return self.encode(x.permute(0, 1, 3, 2)), None

    def encode(self, spec: torch.Tensor) -> torch.Tensor:
        """spec (B, S, F=128, T=66) normalised log-mel -> (B, S, 6, 768) fp32."""
        ops.require_cuda(spec, 'spec')
        if spec.dim() != 4 or tuple(spec.shape[2:]) != (128, 66):
            raise ValueError(f'expected spectrogram of shape (B, S, 128, 66), got {tuple(spec.shape)}')
        B, S = spec.shape[:2]
        n = B * S
        if _tower_trains(self):
            return train_encoders.ast_features(self, spec)
        P, W = self.weights()
        e = 'ast.embeddings.'
        a = ops.im2col_ast(spec.float().contiguous().view(n, 128, 66))
        patch = ops.gemm(a, W['pe_w'], P[e + 'patch_embeddings.projection.bias'], out_f32=True)
        x = ops.ast_tokens(patch, P[e + 'cls_token'], P[e + 'distillation_token'], P[e + 'position_embeddings'], n)   # (n*74, 768) fp32
        self._tap('a_embed', x, A_TOK)
        M = n * A_TOK
        dev = x.device
        ln = torch.empty((M, D), device=dev, dtype=torch.bfloat16)
        qkv = torch.empty((M, 3 * D), device=dev, dtype=torch.bfloat16)
        att = torch.empty((M, D), device=dev, dtype=torch.bfloat16)
        hid = torch.empty((M, 4 * D), device=dev, dtype=torch.bfloat16)
        row, seg = 3 * D, A_TOK * 3 * D
        for i in range(12):                                                                    # ASTLayer modeling_ast.py:294-322
            l = f'ast.encoder.layer.{i}.'
            ops.layernorm(x, P[l + 'layernorm_before.weight'], P[l + 'layernorm_before.bias'], EPS_A, out=ln)
            ops.gemm(ln, W[l + 'qkv'], W[l + 'qkv_b'], out=qkv)
            ops.attention(qkv, qkv[:, D:], qkv[:, 2 * D:], att, q_strides=(seg, 0, row), kv_strides=(seg, 0, row), o_strides=(A_TOK * D, 0, D),
                          n_outer=n, n_inner=1, n_heads=12, head_dim=64, Lq=A_TOK, Lk=A_TOK, scale=0.125)
            ops.gemm(att, W[l + 'o'], P[l + 'attention.output.dense.bias'], out=x, residual=x, out_f32=True)
            ops.layernorm(x, P[l + 'layernorm_after.weight'], P[l + 'layernorm_after.bias'], EPS_A, out=ln)
            ops.gemm(ln, W[l + 'fc1'], P[l + 'intermediate.dense.bias'], out=hid, gelu=True)
            ops.gemm(hid, W[l + 'fc2'], P[l + 'output.dense.bias'], out=x, residual=x, out_f32=True)
        # final layernorm (modeling_ast.py:543) on the 72 patch tokens (ast.py:232-233) fused with the aggregator's norm1;
        # rows stay in (seg, f, t) order and the aggregator walks them with stride 6 (ast.py:266-268 without the permute copy)
        g = 'freq_attn_agg.'
        if self._taps is not None:          # the product path never materialises the un-gathered final norm; the tap computes it on the side
            self._tap('a_last_hidden', ops.layernorm(x, P['ast.layernorm.weight'], P['ast.layernorm.bias'], EPS_A, out_f32=True), A_TOK)
        kv_src = ops.layernorm(x, P['ast.layernorm.weight'], P['ast.layernorm.bias'], EPS_A, out=ln[:n * 72], rows=n * 72, group=72,
                               group_stride=A_TOK, offset=2, gamma2=P[g + 'norm1.weight'], beta2=P[g + 'norm1.bias'], eps2=EPS_V)
        feats = _cls_aggregator(P, W, g, kv_src, n, A_T, 72, 1, A_T, A_F).view(B, S, A_T, D)
        return _MeanTokens.apply(feats) if self.time_pool else feats


class GlobalTransformer(_KernelModule):
    """Synchronisation transformer (sync_model.py:117-173): 3 pre-norm blocks, 8 heads x 96, LN eps 1e-5, erf GELU.
    `pos_emb_cfg.pos_emb` is the RandInitPositionalEncoding table (modules/transformer.py:120-130)."""
    _HEAD = 'off_head'

    def __init__(self, tok_pdrop, embd_pdrop, resid_pdrop, attn_pdrop, n_layer, n_head, n_embd, pos_emb_cfg=None, off_head_cfg=None):
        super().__init__()
        if (n_layer, n_head, n_embd) != (3, 8, D):
            raise NotImplementedError('GlobalTransformer kernels are built for n_layer=3, n_head=8, n_embd=768 (configs/sync.yaml:41-45)')
        if pos_emb_cfg is None or 'RandInitPositionalEncoding' not in pos_emb_cfg['target']:
            raise NotImplementedError('pos_emb_cfg must be RandInitPositionalEncoding (configs/sync.yaml:50-54)')
        self.n_layer, self.n_head, self.n_embd = n_layer, n_head, n_embd
        self.tok_pdrop, self.embd_pdrop, self.resid_pdrop, self.attn_pdrop = tok_pdrop, embd_pdrop, resid_pdrop, attn_pdrop
        block_shape = [int(b) for b in pos_emb_cfg['params']['block_shape']]
        if len(block_shape) != 1:
            raise NotImplementedError('block_shape must be [sequence_length]')
        schema = {k[len('transformer.'):]: v for k, v in state_dict_schema(1, self._n_out(off_head_cfg), self._HEAD).items()
                  if k.startswith('transformer.')}
        schema['pos_emb_cfg.pos_emb'] = (1, block_shape[0], n_embd)
        _build_tree(self, schema)
        _init_reference_like(self)

    @staticmethod
    def _n_out(off_head_cfg):
        if off_head_cfg is None:
            raise NotImplementedError('off_head_cfg is required (configs/sync.yaml:55-59)')
        return int(off_head_cfg['params']['out_features'])

    def _pack(self, P):
        W = {}
        for i in range(3):
            b = f'blocks.{i}.'
            W[b + 'qkv'] = self._bf16(torch.cat([P[b + 'attn.query.weight'], P[b + 'attn.key.weight'], P[b + 'attn.value.weight']], dim=0))
            W[b + 'qkv_b'] = torch.cat([P[b + 'attn.query.bias'], P[b + 'attn.key.bias'], P[b + 'attn.value.bias']], dim=0).contiguous()
            W[b + 'proj'] = self._bf16(P[b + 'attn.proj.weight'])
            W[b + 'fc1'] = self._bf16(P[b + 'mlp.0.weight'])
            W[b + 'fc2'] = self._bf16(P[b + 'mlp.2.weight'])
        return W

    def forward(self, v: torch.Tensor, a: torch.Tensor, targets=None, attempt_to_apply_heads=True):
        """v (B, 8S, 768), a (B, 6S, 768) projected features -> logits (B, n_cls).  sync_model.py:150-173 (eval: dropouts are identity)."""
        ops.require_cuda(v, 'v')
        if self.training:
            # dropout + autograd-visible forward / hand-written backward (train.py).  eval() always takes the inference
            # path below (no autograd graph), whatever torch.is_grad_enabled() says.
            if not attempt_to_apply_heads and self._HEAD == 'off_head':
                raise NotImplementedError('training without the classification head is not implemented')
            return train.sync_transformer(self, v, a)
        B, Sv, _ = v.shape
        Sa = a.shape[1]
        if Sv % 8 or Sa % 6 or Sv // 8 != Sa // 6:
            raise ValueError(f'expected 8 visual and 6 audio tokens per segment, got {Sv} and {Sa}')
        S = Sv // 8
        T = 2 + 14 * S
        P, W = self.weights()
        pos = P['pos_emb_cfg.pos_emb']
        if pos.shape[1] != T:                                    # transformer.py:129-130 adds the full table without slicing
            raise RuntimeError(f'pos_emb has {pos.shape[1]} positions but the sequence has {T}; set block_shape=[{T}]')
        x = ops.sync_tokens(v.float().contiguous(), a.float().contiguous(), P['vis_in_lnorm.weight'], P['vis_in_lnorm.bias'], P['aud_in_lnorm.weight'],
                            P['aud_in_lnorm.bias'], EPS_S, P['OFF_tok'], P['MOD_tok'], pos, B, S)                # (B*T, 768) fp32
        M = B * T
        dev = x.device
        ln = torch.empty((M, D), device=dev, dtype=torch.bfloat16)
        qkv = torch.empty((M, 3 * D), device=dev, dtype=torch.bfloat16)
        att = torch.empty((M, D), device=dev, dtype=torch.bfloat16)
        hid = torch.empty((M, 4 * D), device=dev, dtype=torch.bfloat16)
        row = 3 * D
        for i in range(3):                                                                     # Block.forward transformer.py:94-97
            b = f'blocks.{i}.'
            ops.layernorm(x, P[b + 'ln1.weight'], P[b + 'ln1.bias'], EPS_S, out=ln)
            ops.gemm(ln, W[b + 'qkv'], W[b + 'qkv_b'], out=qkv)
            ops.attention(qkv, qkv[:, D:], qkv[:, 2 * D:], att, q_strides=(T * row, 0, row), kv_strides=(T * row, 0, row), o_strides=(T * D, 0, D),
                          n_outer=B, n_inner=1, n_heads=8, head_dim=96, Lq=T, Lk=T, scale=1.0 / math.sqrt(96.0))
            ops.gemm(att, W[b + 'proj'], P[b + 'attn.proj.bias'], out=x, residual=x, out_f32=True)
            ops.layernorm(x, P[b + 'ln2.weight'], P[b + 'ln2.bias'], EPS_S, out=ln)
            ops.gemm(ln, W[b + 'fc1'], P[b + 'mlp.0.bias'], out=hid, gelu=True)
            ops.gemm(hid, W[b + 'fc2'], P[b + 'mlp.2.bias'], out=x, residual=x, out_f32=True)
        head = self._HEAD if attempt_to_apply_heads or self._HEAD != 'off_head' else None
        if head is None:
            return ops.layernorm(x, P['ln_f.weight'], P['ln_f.bias'], EPS_S, out_f32=True).view(B, T, D)
        return ops.sync_head(x, T, P['ln_f.weight'], P['ln_f.bias'], EPS_S, P[head + '.weight'], P[head + '.bias'], B)


class GlobalTransformerWithSyncabilityHead(GlobalTransformer):
    """sync_model.py:176-190: same stem, 2-class `sync_head` instead of `off_head` (configs/ft_synchability.yaml:41-55)."""
    _HEAD = 'sync_head'

    @staticmethod
    def _n_out(off_head_cfg):
        return 2


_CLASS_BY_NAME = {'MotionFormer': MotionFormer, 'AST': AST, 'GlobalTransformer': GlobalTransformer,
                  'GlobalTransformerWithSyncabilityHead': GlobalTransformerWithSyncabilityHead, 'Linear': nn.Linear}


def instantiate_from_config(config: Mapping[str, Any]) -> nn.Module:
    """utils/utils.py:85-88 semantics, resolving the reference's dotted targets to the classes of this package
    (the class name is what matters: 'model.modules.feat_extractors.audio.ast.AST' -> synchformer_b200.model.AST)."""
    if 'target' not in config:
        raise KeyError('Expected key `target` to instantiate.')
    name = str(config['target']).rsplit('.', 1)[-1]
    if name not in _CLASS_BY_NAME:
        raise NotImplementedError(f"target {config['target']} is outside the B200 hot path (supported: {sorted(_CLASS_BY_NAME)})")
    params = config.get('params', dict()) or dict()
    return _CLASS_BY_NAME[name](**{k: v for k, v in params.items()})


class Synchformer(nn.Module):
    """Drop-in for model.sync_model.Synchformer (sync_model.py:23-114)."""

    def __init__(self, afeat_extractor, vfeat_extractor, aproj, vproj, transformer):
        super().__init__()
        self.vfeat_extractor = instantiate_from_config(vfeat_extractor)
        self.afeat_extractor = instantiate_from_config(afeat_extractor)
        self.vproj = instantiate_from_config(vproj)
        self.aproj = instantiate_from_config(aproj)
        self.transformer = instantiate_from_config(transformer)
        if not isinstance(self.vproj, nn.Linear) or not isinstance(self.aproj, nn.Linear):
            raise NotImplementedError('vproj / aproj must be torch.nn.Linear (configs/sync.yaml:28-39)')
        with torch.no_grad():
            for lin in (self.vproj, self.aproj):                                    # init_weights sync_model.py:13-20 is applied by the
                lin.weight.normal_(0.0, 0.02)                                       # reference only inside GlobalTransformer; harmless here
                lin.bias.zero_()
        self._proj_cache, self._proj_key = {}, None

    # ---- helpers -------------------------------------------------------------------------------------------------
    def _proj_weights(self):
        ps = (self.vproj.weight, self.aproj.weight)
        key = tuple((p.device, p._version, p.data_ptr()) for p in ps)
        if key != self._proj_key:
            with torch.no_grad():
                self._proj_cache = {'v': ops.cast_bf16(self.vproj.weight.detach().contiguous()),
                                    'a': ops.cast_bf16(self.aproj.weight.detach().contiguous())}
            self._proj_key = key
        return self._proj_cache

    def project(self, vis: torch.Tensor, aud: torch.Tensor, out_v: Optional[torch.Tensor] = None, out_a: Optional[torch.Tensor] = None):
        """vproj / aproj (sync_model.py:55-56) + segment flattening (:59-62).  (B,S,8,768),(B,S,6,768) -> (B,8S,768),(B,6S,768) fp32.
        out_v (B*S*8, 768) / out_a (B*S*6, 768) fp32: inference only - the GEMM epilogues write there (parallel.py passes slices of the
        all-gather send buffer, so no copy sits between the projection and the collective)."""
        B, S = vis.shape[:2]
        Wp = self._proj_weights()
        if self.training:
            if out_v is not None or out_a is not None:
                raise NotImplementedError('project(out_v=, out_a=) is an inference-path option')
            v = train.linear(vis.float().contiguous().view(-1, D), self.vproj.weight, self.vproj.bias, Wp['v'])        # N3: differentiable
            a = train.linear(aud.float().contiguous().view(-1, D), self.aproj.weight, self.aproj.bias, Wp['a'])
            return v.view(B, S * 8, D), a.view(B, S * 6, D)
        v = ops.gemm(ops.cast_bf16(vis.float().contiguous().view(-1, D)), Wp['v'], self.vproj.bias.detach(), out=out_v, out_f32=True)
        a = ops.gemm(ops.cast_bf16(aud.float().contiguous().view(-1, D)), Wp['a'], self.aproj.bias.detach(), out=out_a, out_f32=True)
        return v.view(B, S * 8, D), a.view(B, S * 6, D)

    # ---- reference API -------------------------------------------------------------------------------------------
    def forward(self, vis: torch.Tensor, aud: torch.Tensor, targets: torch.Tensor = None, for_loop=False, vis_mask: torch.Tensor = None,
                aud_mask: torch.Tensor = None, loss_fn=None):
        """vis (B, S, Tv=16, C=3, H=224, W=224), aud (B, S, 1, F=128, Ta=66) -> (loss | None, logits (B, n_cls))."""
        vis = self.extract_vfeats(vis, for_loop, vis_mask=vis_mask)
        aud = self.extract_afeats(aud, for_loop, aud_mask=aud_mask)
        v, a = self.project(vis, aud)
        logits = self.transformer(v, a)
        loss = self.compute_loss(logits, targets, loss_fn)
        return loss, logits

    @staticmethod
    def segment_ranges(n_vframes: int, n_aframes: int, n_segments: int, segment_size_vframes: int = 16, step_size_seg: float = 0.5,
                       v_fps: int = 25, a_fps: int = 16000):
        """(v_start, v_stride, a_start, a_stride) of GenerateMultipleSegments with is_start_random=False
        (dataset/transforms.py:421-499; configs/sync.yaml:93-101, 170-176): the segment sequence is centred in the clip."""
        seg_a = int(segment_size_vframes / v_fps * a_fps)
        v_stride, a_stride = int(step_size_seg * segment_size_vframes), int(step_size_seg * seg_a)
        seq = n_segments * step_size_seg + (1 - step_size_seg)
        v_len = int(seq * segment_size_vframes)
        if v_len > n_vframes or int(seq * seg_a) > n_aframes:
            raise ValueError(f'cant make {n_segments} segs of len {segment_size_vframes} in a vid of len {n_vframes}')
        v_start = (n_vframes - v_len) // 2
        a_start = int(v_start / v_fps * a_fps)
        return v_start, v_stride, a_start, a_stride

    @torch.no_grad()
    def forward_clip(self, frames: torch.Tensor, waveform: torch.Tensor, n_segments: int = 14, **seg_kwargs) -> torch.Tensor:
        """un-segmented inputs straight from the decoder - frames (B, n_frames, 3, 224, 224) uint8 / fp16 / fp32 and
        waveform (B, n_samples) fp32 at 16 kHz -> offset logits (B, n_cls).  Segment slicing, uint8 normalisation and the mel front-end
        all happen inside the kernels; equivalent to the transform chain + forward() on the explicit (B, S, ...) tensors."""
        v0, vs, a0, as_ = self.segment_ranges(frames.shape[1], waveform.shape[1], n_segments, **seg_kwargs)
        vis = self.vfeat_extractor.encode_clip(frames, n_segments, v0, vs)
        mel = ops.mel_frontend_clip(waveform.float().contiguous(), n_segments, a0, as_)
        aud = self.afeat_extractor.encode(mel)
        v, a = self.project(vis, aud)
        return self.transformer(v, a)

    def extract_vfeats(self, vis, for_loop=False, vis_mask=None):
        if vis_mask is not None:
            raise NotImplementedError('vis_mask is not supported (no caller in the reference passes it)')
        if _tower_trains(self.vfeat_extractor):               # is_trainable: True -> gradients flow into the tower (train_encoders.py)
            return self.vfeat_extractor.encode(vis)
        with torch.no_grad():
            return self.vfeat_extractor.encode(vis)           # for_loop only trades memory for speed in the reference; results are identical

    def extract_afeats(self, aud, for_loop=False, aud_mask=None):
        if aud_mask is not None:
            raise NotImplementedError('aud_mask is not supported (no caller in the reference passes it)')
        B, S, _, Fa, Ta = aud.shape
        if _tower_trains(self.afeat_extractor):
            return self.afeat_extractor.encode(aud.view(B, S, Fa, Ta))
        with torch.no_grad():
            return self.afeat_extractor.encode(aud.view(B, S, Fa, Ta))

    def compute_loss(self, logits, targets, loss_fn: str = None):
        loss = None
        if targets is not None:
            if loss_fn is None or loss_fn == 'cross_entropy':
                if logits.is_cuda and logits.requires_grad and targets.dim() == 1 and targets.dtype == torch.int64:
                    from . import optim                          # training step: loss and d loss / d logits in one launch
                    loss = optim.cross_entropy(logits, targets)
                else:
                    loss = torch.nn.functional.cross_entropy(logits, targets)
            else:
                raise NotImplementedError(f'Loss {loss_fn} not implemented')
        return loss

    def load_state_dict(self, sd: Mapping[str, Any], strict: bool = True):
        """Trims a longer checkpoint pos-emb, rejects a shorter one (sync_model.py:101-114)."""
        key = 'transformer.pos_emb_cfg.pos_emb'
        if key in sd:
            weight_len = sd[key].shape[1]
            self_len = self.transformer.pos_emb_cfg.pos_emb.shape[1]
            if weight_len > self_len:
                sd = dict(sd)
                sd[key] = sd[key][:, :self_len, :]
                logging.warning(f'Trimming the state dict for pos emb from {weight_len} to {self_len}')
            elif weight_len < self_len:
                raise ValueError(f'Cant load state dict with shorter seq len ({weight_len} vs {self_len})')
        return super().load_state_dict(sd, strict)


class GraphedForward:
    """CUDA-graph replay of `Synchformer.forward` for one fixed input shape (small-batch latency: the forward is ~290 kernel launches,
    which at B = 1 are launch-bound when issued one by one from Python).  Inputs are copied into static device buffers, the captured
    graph is replayed on the current stream, and the static logits tensor is returned (valid until the next call).

        fwd = GraphedForward(model, vis_example, aud_example)
        logits = fwd(vis, aud)          # same shapes / dtypes as the examples
    """

    def __init__(self, model: 'Synchformer', vis: torch.Tensor, aud: torch.Tensor, warmup: int = 2):
        ops.require_cuda(vis, 'vis')
        self.model = model
        self.vis = torch.empty_like(vis)
        self.aud = torch.empty_like(aud)
        self.vis.copy_(vis)
        self.aud.copy_(aud)
        stream = torch.cuda.Stream(device=vis.device)
        stream.wait_stream(torch.cuda.current_stream(vis.device))
        with torch.cuda.stream(stream), torch.no_grad():
            for _ in range(max(1, warmup)):                     # weight caches, function attributes, constant tables: all set up before capture
                model(self.vis, self.aud)
        torch.cuda.current_stream(vis.device).wait_stream(stream)
        torch.cuda.synchronize(vis.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            _, self.logits = model(self.vis, self.aud)

    def __call__(self, vis: torch.Tensor, aud: torch.Tensor) -> torch.Tensor:
        if vis.shape != self.vis.shape or aud.shape != self.aud.shape or vis.dtype != self.vis.dtype or aud.dtype != self.aud.dtype:
            raise ValueError(f'graph was captured for {tuple(self.vis.shape)} {self.vis.dtype} / {tuple(self.aud.shape)} {self.aud.dtype}')
        self.vis.copy_(vis, non_blocking=True)
        self.aud.copy_(aud, non_blocking=True)
        self.graph.replay()
        return self.logits


def sync_yaml_model_config(n_segments: int = 14, n_classes: int = 21, transformer_target: str = 'model.sync_model.GlobalTransformer') -> dict:
    """The `model.params` tree of configs/sync.yaml:6-59 with interpolations resolved, as plain dicts."""
    return dict(
        afeat_extractor=dict(is_trainable=False, target='model.modules.feat_extractors.audio.ast.AST',
                             params=dict(ckpt_path=None, extract_features=True, max_spec_t=66, factorize_freq_time=True,
                                         agg_freq_module='TransformerEncoderLayer', agg_time_module='torch.nn.Identity', add_global_repr=False)),
        vfeat_extractor=dict(is_trainable=False, target='model.modules.feat_extractors.visual.motionformer.MotionFormer',
                             params=dict(ckpt_path=None, extract_features=True, factorize_space_time=True,
                                         agg_space_module='TransformerEncoderLayer', agg_time_module='torch.nn.Identity', add_global_repr=False)),
        aproj=dict(target='torch.nn.Linear', params=dict(in_features=768, out_features=768)),
        vproj=dict(target='torch.nn.Linear', params=dict(in_features=768, out_features=768)),
        transformer=dict(target=transformer_target,
                         params=dict(n_layer=3, n_head=8, n_embd=768, tok_pdrop=0.0, embd_pdrop=0.1, resid_pdrop=0.1, attn_pdrop=0.1,
                                     pos_emb_cfg=dict(target='model.modules.transformer.RandInitPositionalEncoding',
                                                      params=dict(block_shape=[2 + 14 * n_segments], n_embd=768)),
                                     off_head_cfg=dict(target='torch.nn.Linear', params=dict(in_features=768, out_features=n_classes)))),
    )


def build_synchformer(n_segments: int = 14, n_classes: int = 21, state_dict: Optional[Mapping[str, torch.Tensor]] = None, device=None) -> Synchformer:
    """Convenience constructor for tests / bench: sync.yaml architecture, optional weights, eval mode."""
    cfg = sync_yaml_model_config(n_segments, n_classes)
    cfg = {k: {kk: vv for kk, vv in v.items() if kk != 'is_trainable'} for k, v in cfg.items()}
    model = Synchformer(**cfg)
    if state_dict is not None:
        model.load_state_dict(state_dict, strict=True)
    model.eval()
    if device is not None:
        model.to(device)
    return model
