"""Stage-I contrastive forward around the same two encoders (BASELINE.json config 3).

Mirrors `AVCLIP.forward` of the reference (model/modules/feat_extractors/train_clip_src/open_clip/model.py:449-545) for the
forward pass in eval / no_grad: encoders with `agg_time_module='AveragePooling'` -> (B*S, 768) -> identity bridges
(`DoNothingBridge`, configs/segment_avclip.yaml:45-54) -> L2 normalise -> similarities / logit_scale -> symmetric cross-entropy with
identity targets.  Inputs come in the STAGE-I layouts: rgb (B, S, C=3, T=16, H, W), audio (B, S, T=66, F=128)
(segment_avclip.yaml:208-211).  The encoders are the kernel-backed `MotionFormer` / `AST` of model.py; the tail (L2 normalise, similarities, symmetric
cross-entropy and their backward) runs on the kernels of csrc/contrastive.cu; with `gather_for_loss=True` the normalised features of all
ranks are all-gathered with an autograd-aware collective (open_clip/model.py:492-494), torch.distributed being the plumbing.
In train mode (with autograd on) the towers take their differentiable path (train_encoders.py, SURVEY.md §8f N1: DropPath + hand-written
backward), so `out['losses']['segment_contrastive_loss'].backward()` fills the gradients of both encoders and of `logit_scale`.
"""
import logging

import torch
import torch.distributed as dist
from torch import nn

from . import ops
from .model import AST, MotionFormer


class _L2Normalize(torch.autograd.Function):
    """F.normalize(x, dim=-1) (open_clip/model.py:543) on (n, D) fp32."""

    @staticmethod
    def forward(ctx, x):
        xn, inv = ops.l2_normalize(x.float().contiguous())
        ctx.save_for_backward(xn, inv)
        return xn

    @staticmethod
    def backward(ctx, g):
        xn, inv = ctx.saved_tensors
        return ops.l2_normalize_bwd(xn, inv, g.float().contiguous())


class _AllGatherRows(torch.autograd.Function):
    """torch.distributed.nn.all_gather + cat(dim=0) (open_clip/model.py:492-494): forward all-gathers the (n, D) blocks of every rank,
    backward returns this rank's block of the SUM over ranks of the gathered gradient (reduce-scatter)."""

    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        world = dist.get_world_size(group)
        out = torch.empty((world * x.shape[0],) + tuple(x.shape[1:]), device=x.device, dtype=x.dtype)
        dist.all_gather_into_tensor(out, x.contiguous(), group=group)
        return out

    @staticmethod
    def backward(ctx, g):
        world, rank = dist.get_world_size(ctx.group), dist.get_rank(ctx.group)
        g = g.contiguous()
        n = g.shape[0] // world
        if dist.get_backend(ctx.group) == 'gloo':            # gloo has no reduce-scatter: all-reduce and keep this rank's block
            dist.all_reduce(g, group=ctx.group)
            return g[rank * n:(rank + 1) * n].clone(), None
        out = torch.empty((n,) + tuple(g.shape[1:]), device=g.device, dtype=g.dtype)
        dist.reduce_scatter_tensor(out, g, group=ctx.group)
        return out, None


class _ContrastiveLoss(torch.autograd.Function):
    """compute_loss (open_clip/model.py:507-527): one launch pair forward (similarities, softmax, loss, d loss / d sim), one backward.
    vn_all / an_all None = no gathering: the keys are the local rows."""

    @staticmethod
    def forward(ctx, vn, an, vn_all, an_all, logit_scale):
        local = vn_all is None
        scale = logit_scale.detach().float().reshape(1)
        loss, dscale, G = ops.contrastive_loss(vn, an, vn if local else vn_all, an if local else an_all, scale)
        ctx.local = local
        ctx.save_for_backward(vn, an, *(() if local else (vn_all, an_all)), G, scale, dscale)
        return loss.view(())

    @staticmethod
    def backward(ctx, g):
        saved = ctx.saved_tensors
        vn, an = saved[0], saved[1]
        vn_all, an_all = (None, None) if ctx.local else (saved[2], saved[3])
        G, scale, dscale = saved[-3:]
        d_vn, d_an, d_vn_all, d_an_all, d_scale = ops.contrastive_loss_bwd(vn, an, vn_all, an_all, G, scale, g.float().reshape(1).contiguous(), dscale)
        return d_vn, d_an, d_vn_all, d_an_all, d_scale.view(())


class DoNothingBridge(nn.Identity):
    """model.modules.bridges.DoNothingBridge (configs/segment_avclip.yaml:45-54): identity that accepts in_features / out_features."""

    def __init__(self, in_features: int = None, out_features: int = None, **kwargs):
        super().__init__()


def _tower_from_config(cfg, default_cls):
    """`{target, params[, is_trainable]}` (segment_avclip.yaml:12-44) -> kernel-backed tower.  A stage-I `.pt` `ckpt_path` initialises the
    tower (model._init_from_stage1_ckpt); the downloadable pre-trained initialisations (HF hub name / SSv2 .pyth) are dropped with a warning."""
    if cfg is None:
        return None
    params = dict(cfg.get('params', {}) or {})
    if params.get('ckpt_path') is not None and not str(params['ckpt_path']).endswith('.pt'):
        # the public pre-trained initialisations (HF AudioSet AST, SSv2 Motionformer .pyth) are downloads: no network here
        logging.warning(f"synchformer_b200.AVCLIP: ckpt_path={params['ckpt_path']!r} is ignored - initialise the towers with load_state_dict")
        params.pop('ckpt_path')
    name = str(cfg['target']).rsplit('.', 1)[-1]
    if name != default_cls.__name__:
        raise NotImplementedError(f"tower target {cfg['target']} is outside the B200 hot path (expected ...{default_cls.__name__})")
    tower = default_cls(**params)
    if cfg.get('is_trainable', True) is False:
        tower.requires_grad_(False)
    return tower


class AVCLIP(nn.Module):
    """Same constructor keywords as the reference class (open_clip/model.py:451-470), so `model.target=synchformer_b200.avclip.AVCLIP` is the
    only change to configs/segment_avclip.yaml; with no tower configs it builds the segment_avclip.yaml towers."""

    def __init__(self, n_embd: int = 768, afeat_extractor=None, vfeat_extractor=None, aproj=None, vproj=None, init_scale: float = 0.07,
                 clamp_scale_min: float = 0.001, clamp_scale_max: float = 0.5, gather_for_loss: bool = False):
        super().__init__()
        self.output_dict = True
        self.n_embd = n_embd
        self.v_encoder = _tower_from_config(vfeat_extractor, MotionFormer) or MotionFormer(
            extract_features=True, factorize_space_time=True, agg_space_module='TransformerEncoderLayer', agg_time_module='AveragePooling',
            add_global_repr=False)
        self.a_encoder = _tower_from_config(afeat_extractor, AST) or AST(
            extract_features=True, max_spec_t=66, factorize_freq_time=True, agg_freq_module='TransformerEncoderLayer',
            agg_time_module='AveragePooling', add_global_repr=False)
        for cfg in (aproj, vproj):
            if cfg is not None and 'DoNothingBridge' not in str(cfg['target']):
                raise NotImplementedError(f"bridge {cfg['target']}: only DoNothingBridge (segment_avclip.yaml:45-54) is implemented")
        self.vproj, self.aproj = DoNothingBridge(), DoNothingBridge()
        self.clamp_scale_min, self.clamp_scale_max, self.init_scale, self.gather_for_loss = clamp_scale_min, clamp_scale_max, init_scale, bool(gather_for_loss)
        self.logit_scale = nn.Parameter(torch.ones([]) * init_scale)

    @torch.no_grad()
    def clamp_logit_scales(self):
        """open_clip/model.py:579-582: clamps the learned temperature in place (called at the top of every forward, :490)."""
        self.logit_scale.clamp_(self.clamp_scale_min, self.clamp_scale_max)
        return (self.logit_scale, None)

    def encode_streams(self, vis: torch.Tensor, aud: torch.Tensor, do_norm: bool = True):
        """open_clip/model.py:529-545: (B*S, 768) visual and audio segment features."""
        v, _ = self.v_encoder(vis)                 # (B, S, 768)
        a, _ = self.a_encoder(aud)
        v, a = v.reshape(-1, self.n_embd), a.reshape(-1, self.n_embd)
        if do_norm:
            v, a = _L2Normalize.apply(v), _L2Normalize.apply(a)
        return v, a

    def forward(self, vis: torch.Tensor, aud: torch.Tensor, alpha: float = 0.0, for_loop: bool = False, world_size: int = 1):
        assert alpha == 0.0, f'alpha={alpha} not supported yet'          # same assertion as the reference (:489)
        with torch.set_grad_enabled(torch.is_grad_enabled() and self.training):     # eval: inference kernels, nothing is recorded
            return self._forward(vis, aud, world_size)

    def _forward(self, vis: torch.Tensor, aud: torch.Tensor, world_size: int = 1):
        logit_scales = self.clamp_logit_scales()
        vfeat, afeat = self.encode_streams(vis, aud)
        if world_size > 1 and self.gather_for_loss:                       # :492-494 global negatives; targets stay eye(n, N) as in the reference
            vfeat_all, afeat_all = _AllGatherRows.apply(vfeat, None), _AllGatherRows.apply(afeat, None)
        else:
            vfeat_all, afeat_all = None, None
        loss = _ContrastiveLoss.apply(vfeat, afeat, vfeat_all, afeat_all, self.logit_scale)          # compute_loss :507-527
        return {'rgb_features': (vfeat, None), 'audio_features': (afeat, None), 'logit_scales': logit_scales,
                'losses': {'segment_contrastive_loss': loss}}


def shift_and_get_preds(a: torch.Tensor, v: torch.Tensor, W: int):
    """Zero-shot shifted-window evaluation of the stage-I features (training/train.py:549-579): for every window of W
    consecutive segments of the audio track, the most similar window of the visual track and vice versa.
    a, v (B, S, D) segment features -> (preds_a, preds_v), each (B, S - W + 1) int64.

    sim[b, i, j] = <a[b, i:i+W], v[b, j:j+W]> = sum_w seg[b, i + w, j + w] with seg[b] = a[b] v[b]^T: the segment-by-segment products of the
    whole batch are ONE tensor-core GEMM ((B S, D) x (B S, D)^T, fp32 out; the cross-clip blocks are discarded), the window sums and the
    arg-max are index plumbing on the (B, S, S) result."""
    from . import ops
    assert a.shape == v.shape and a.dim() == 3, f'{tuple(a.shape)} != {tuple(v.shape)}'
    B, S, D = a.shape
    n_shifts = S - W + 1
    assert n_shifts >= 1, f'window {W} longer than the {S} segments'
    ops.require_cuda(a, 'a')
    ab = ops.cast_bf16(a.float().contiguous().view(B * S, D))
    vb = ops.cast_bf16(v.float().contiguous().view(B * S, D))
    pad = (-(B * S)) % 8                                         # the GEMM wants N % 8 == 0: pad the visual rows with zeros
    if pad:
        vb = torch.cat([vb, vb.new_zeros((pad, D))], dim=0)
    full = ops.gemm(ab, vb, None, out_f32=True)[:, :B * S].reshape(B, S, B, S)
    idx = torch.arange(B, device=a.device)
    seg = full[idx, :, idx, :]                                   # (B, S, S): a[b, i] . v[b, j]
    win = seg.unfold(1, W, 1).unfold(2, W, 1)                    # (B, n, n, W, W)
    sim = win.diagonal(dim1=-2, dim2=-1).sum(-1)                 # (B, n, n)
    return torch.argmax(sim, dim=-2), torch.argmax(sim, dim=-1)
