"""training step of the synchronisation module on the sm_100a kernels.

What `scripts/train_sync.py:177-183` + `scripts/train_utils.py:373-386` exercise with the extractors frozen (configs/sync.yaml:8,20):
`loss, logits = model(vid, aud, targets)` in train mode, then `scaler.scale(loss).backward()`.  Only `vproj`, `aproj` and
`transformer.*` (22.6 M parameters) receive gradients.  This module supplies the two `torch.autograd.Function`s that put the CUDA
forward-with-dropout and the hand-written backward behind autograd, so optimisers, `GradScaler`, gradient clipping and
`DistributedDataParallel` (which hooks the parameters' gradient accumulators) work unchanged:

  linear(x, weight, bias)                      vproj / aproj            (sync_model.py:55-56)
  sync_transformer(transformer, v, a)          GlobalTransformer.forward (sync_model.py:150-173) with Block / SelfAttention
                                               (modules/transformer.py:58-97) in training mode

Numerics follow the reference under `torch.autocast(bf16)`: bf16 GEMM operands with fp32 accumulation, fp32 LayerNorm / softmax /
residual stream, GELU on the bf16 pre-activation.  Dropout masks are counter-based (csrc/philox.cuh): a pure function of
(seed, site, element index), regenerated in the backward instead of stored; the seed of a step is drawn from torch's CPU generator
(so `torch.manual_seed` makes a run reproducible), the stream itself differs from torch's CUDA dropout (as it would between any two
torch versions).  Linear backward: dX = dY W and dW = dY^T X run on the same tcgen05 GEMM as the forward, on operands transposed by
`sfb_transpose_bf16`; bias gradients are deterministic column sums.

PyTorch is used for memory, streams and the autograd graph only; there is no eager fallback.
"""
import math
from typing import Dict, List, Optional

import torch

from . import ops

D = 768
EPS_S = 1e-5
N_HEAD, HEAD_DIM = 8, 96

SITE_EMBD = 0


def site_attn(i: int) -> int:
    return 1 + 3 * i


def site_resid_attn(i: int) -> int:
    return 2 + 3 * i


def site_resid_mlp(i: int) -> int:
    return 3 + 3 * i


def draw_seed() -> int:
    """Per-step dropout seed from torch's default CPU generator (no device sync)."""
    return int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())


def _t(x: torch.Tensor) -> torch.Tensor:
    """transposed GEMM operand; a reduction dimension shorter than one 64-wide k-block is zero-padded up to it"""
    return ops.transpose_bf16(x, 64 if x.shape[0] < 64 else 8)


class _LinearFn(torch.autograd.Function):
    """y = x W^T + b  with fp32 x / y and bf16 tensor-core operands (the autocast behaviour of nn.Linear)."""

    @staticmethod
    def forward(ctx, x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, w_bf16: torch.Tensor):
        xb = ops.cast_bf16(x)
        ctx.save_for_backward(xb, w_bf16)
        return ops.gemm(xb, w_bf16, bias.detach(), out_f32=True)

    @staticmethod
    def backward(ctx, dy: torch.Tensor):
        xb, w_bf16 = ctx.saved_tensors
        dy = dy.float().contiguous()
        dyb = ops.cast_bf16(dy)
        dx = ops.gemm(dyb, _t(w_bf16), None, out_f32=True) if ctx.needs_input_grad[0] else None
        dw = ops.gemm(_t(dyb), _t(xb), None, out_f32=True) if ctx.needs_input_grad[1] else None
        db = ops.colsum(dy) if ctx.needs_input_grad[2] else None
        return dx, dw, db, None


def linear(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, w_bf16: torch.Tensor) -> torch.Tensor:
    """x (M, 768) fp32 contiguous -> (M, n_out) fp32, differentiable w.r.t. x, weight and bias."""
    ops.require_cuda(x, 'x')
    return _LinearFn.apply(x, weight, bias, w_bf16)


# ------------------------------------------------------------------------------------------------------------------
# GlobalTransformer: forward with dropout (activations kept for the backward) and the backward itself
# ------------------------------------------------------------------------------------------------------------------
def _forward_train(tr, v: torch.Tensor, a: torch.Tensor, seed: int):
    """tr: model.GlobalTransformer; v (B, 8S, 768), a (B, 6S, 768) fp32 contiguous.  Returns (logits, saved)."""
    P, W = tr.weights()
    B, Sv, _ = v.shape
    S = Sv // 8
    T = 2 + 14 * S
    pos = P['pos_emb_cfg.pos_emb']
    if pos.shape[1] != T:
        raise RuntimeError(f'pos_emb has {pos.shape[1]} positions but the sequence has {T}; set block_shape=[{T}]')
    x = ops.sync_tokens(v, a, P['vis_in_lnorm.weight'], P['vis_in_lnorm.bias'], P['aud_in_lnorm.weight'], P['aud_in_lnorm.bias'], EPS_S,
                        P['OFF_tok'], P['MOD_tok'], pos, B, S)                                                # (B*T, 768) fp32
    if tr.embd_pdrop > 0:
        ops.dropout(x, tr.embd_pdrop, seed, SITE_EMBD, out=x)                                                 # sync_model.py:168
    scale = 1.0 / math.sqrt(HEAD_DIM)
    blocks = []
    for i in range(tr.n_layer):                                                                               # Block.forward transformer.py:94-97
        b = f'blocks.{i}.'
        ln1 = ops.layernorm(x, P[b + 'ln1.weight'], P[b + 'ln1.bias'], EPS_S)
        qkv = ops.gemm(ln1, W[b + 'qkv'], W[b + 'qkv_b'])
        att, lse = ops.attention_train_fwd(qkv, B, T, N_HEAD, HEAD_DIM, scale, tr.attn_pdrop, seed, site_attn(i))
        y = ops.gemm(att, W[b + 'proj'], P[b + 'attn.proj.bias'], out_f32=True)
        x_mid = ops.dropout(y, tr.resid_pdrop, seed, site_resid_attn(i), residual=x)
        ln2 = ops.layernorm(x_mid, P[b + 'ln2.weight'], P[b + 'ln2.bias'], EPS_S)
        hpre = ops.gemm(ln2, W[b + 'fc1'], P[b + 'mlp.0.bias'])
        hid = ops.gelu_fwd(hpre)
        y = ops.gemm(hid, W[b + 'fc2'], P[b + 'mlp.2.bias'], out_f32=True)
        x_out = ops.dropout(y, tr.resid_pdrop, seed, site_resid_mlp(i), residual=x_mid)
        blocks.append((x, ln1, qkv, att, lse, x_mid, ln2, hpre, hid))
        x = x_out
    head = tr._HEAD
    logits = ops.sync_head(x, T, P['ln_f.weight'], P['ln_f.bias'], EPS_S, P[head + '.weight'], P[head + '.bias'], B)
    saved = dict(B=B, S=S, T=T, v=v, a=a, blocks=blocks, x_final=x, seed=seed)
    return logits, saved


def _backward(tr, saved, dlogits: torch.Tensor, need_dv: bool, need_da: bool):
    """Returns (dv | None, da | None, {parameter name: gradient})."""
    P, W = tr.weights()
    B, S, T, seed = saved['B'], saved['S'], saved['T'], saved['seed']
    g: Dict[str, torch.Tensor] = {}
    head = tr._HEAD
    dx, g['ln_f.weight'], g['ln_f.bias'], g[head + '.weight'], g[head + '.bias'] = ops.sync_head_bwd(
        saved['x_final'], T, P['ln_f.weight'], P['ln_f.bias'], EPS_S, P[head + '.weight'], dlogits, B)
    scale = 1.0 / math.sqrt(HEAD_DIM)
    for i in reversed(range(tr.n_layer)):
        b = f'blocks.{i}.'
        x_in, ln1, qkv, att, lse, x_mid, ln2, hpre, hid = saved['blocks'][i]
        # x_out = x_mid + drop(fc2(gelu(fc1(ln2(x_mid)))))                                       transformer.py:96, 86-93
        dy = ops.dropout(dx, tr.resid_pdrop, seed, site_resid_mlp(i), out_bf16=True)            # (M, 768) bf16
        g[b + 'mlp.2.weight'] = ops.gemm(_t(dy), _t(hid), None, out_f32=True)                    # dY^T X  (768, 3072)
        g[b + 'mlp.2.bias'] = ops.colsum(dy)
        dhid = ops.gemm(dy, _t(W[b + 'fc2']), None)                                              # dY W    (M, 3072) bf16
        dpre = ops.gelu_bwd(dhid, hpre)
        g[b + 'mlp.0.weight'] = ops.gemm(_t(dpre), _t(ln2), None, out_f32=True)                  # (3072, 768)
        g[b + 'mlp.0.bias'] = ops.colsum(dpre)
        dln2 = ops.gemm(dpre, _t(W[b + 'fc1']), None, out_f32=True)                              # (M, 768) fp32
        _, g[b + 'ln2.weight'], g[b + 'ln2.bias'] = ops.layernorm_bwd(dln2, x_mid, P[b + 'ln2.weight'], EPS_S, dx=dx, accumulate=True)
        # x_mid = x_in + drop(proj(attention(qkv(ln1(x_in)))))                                   transformer.py:95, 58-76
        dy = ops.dropout(dx, tr.resid_pdrop, seed, site_resid_attn(i), out_bf16=True)
        g[b + 'attn.proj.weight'] = ops.gemm(_t(dy), _t(att), None, out_f32=True)
        g[b + 'attn.proj.bias'] = ops.colsum(dy)
        datt = ops.gemm(dy, _t(W[b + 'proj']), None)                                             # (M, 768) bf16
        dqkv = ops.attention_train_bwd(qkv, att, datt, lse, B, T, N_HEAD, HEAD_DIM, scale, tr.attn_pdrop, seed, site_attn(i))
        dw = ops.gemm(_t(dqkv), _t(ln1), None, out_f32=True)                                     # (2304, 768) = [query; key; value]
        db = ops.colsum(dqkv)
        for k, n in enumerate(('query', 'key', 'value')):
            g[b + f'attn.{n}.weight'] = dw[k * D:(k + 1) * D]
            g[b + f'attn.{n}.bias'] = db[k * D:(k + 1) * D]
        dln1 = ops.gemm(dqkv, _t(W[b + 'qkv']), None, out_f32=True)
        _, g[b + 'ln1.weight'], g[b + 'ln1.bias'] = ops.layernorm_bwd(dln1, x_in, P[b + 'ln1.weight'], EPS_S, dx=dx, accumulate=True)
    # x0 = drop([OFF, LN_v(v), MOD, LN_a(a)] + pos)                                              sync_model.py:153-168
    if tr.embd_pdrop > 0:
        ops.dropout(dx, tr.embd_pdrop, seed, SITE_EMBD, out=dx)
    dpos = ops.colsum(dx.view(B, T * D)).view(T, D)
    g['pos_emb_cfg.pos_emb'] = dpos.view(1, T, D)
    g['OFF_tok'] = dpos[0].clone().view(1, 1, D)
    g['MOD_tok'] = dpos[1 + 8 * S].clone().view(1, 1, D)
    v2, a2 = saved['v'].view(B * 8 * S, D), saved['a'].view(B * 6 * S, D)
    dv, g['vis_in_lnorm.weight'], g['vis_in_lnorm.bias'] = ops.layernorm_bwd(dx, v2, P['vis_in_lnorm.weight'], EPS_S, rows=B * 8 * S,
                                                                               group=8 * S, group_stride=T, offset=1)
    da, g['aud_in_lnorm.weight'], g['aud_in_lnorm.bias'] = ops.layernorm_bwd(dx, a2, P['aud_in_lnorm.weight'], EPS_S, rows=B * 6 * S,
                                                                               group=6 * S, group_stride=T, offset=2 + 8 * S)
    return (dv.view(B, 8 * S, D) if need_dv else None), (da.view(B, 6 * S, D) if need_da else None), g


class _SyncTransformerFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tr, names: List[str], seed: int, v: torch.Tensor, a: torch.Tensor, *params: torch.Tensor):
        logits, saved = _forward_train(tr, v, a, seed)
        ctx.tr, ctx.names, ctx.saved = tr, names, saved
        return logits

    @staticmethod
    def backward(ctx, dlogits: torch.Tensor):
        saved = ctx.saved
        if saved is None:
            raise RuntimeError('backward through the sync transformer a second time: activations were freed after the first pass')
        ctx.saved = None
        dv, da, g = _backward(ctx.tr, saved, dlogits.float().contiguous(), ctx.needs_input_grad[3], ctx.needs_input_grad[4])
        shapes = {n: tuple(p.shape) for n, p in ctx.tr.named_parameters()}
        grads = [g[n].reshape(shapes[n]) if ctx.needs_input_grad[5 + k] else None for k, n in enumerate(ctx.names)]
        return (None, None, None, dv, da, *grads)


def sync_transformer(tr, v: torch.Tensor, a: torch.Tensor, seed: Optional[int] = None) -> torch.Tensor:
    """GlobalTransformer.forward in training mode: v (B, 8S, 768), a (B, 6S, 768) -> logits (B, n_cls), differentiable w.r.t. v, a and
    every parameter of `tr`."""
    ops.require_cuda(v, 'v')
    if tr.tok_pdrop > 0:
        raise NotImplementedError('tok_pdrop > 0 (whole-token Dropout1d, sync_model.py:160-161) is not implemented; configs/sync.yaml:46 uses 0.0')
    B, Sv, _ = v.shape
    Sa = a.shape[1]
    if Sv % 8 or Sa % 6 or Sv // 8 != Sa // 6:
        raise ValueError(f'expected 8 visual and 6 audio tokens per segment, got {Sv} and {Sa}')
    params = dict(tr.named_parameters())
    names = list(params)
    if seed is None:
        seed = draw_seed()
    return _SyncTransformerFn.apply(tr, names, seed, v.float().contiguous(), a.float().contiguous(), *params.values())
