"""CPU stand-ins for the kernels behind `synchformer_b200.ops`, for testing HOST LOGIC without a GPU.

TEST INFRASTRUCTURE ONLY.  Each function restates the documented contract of one C-ABI entry point (include/synchformer_b200.h)
in plain torch on CPU tensors, so that `-m "not gpu"` tests can run the orchestration in `synchformer_b200/train.py` (which saved
tensor feeds which GEMM, operand transposes, gradient slicing, LayerNorm gathers, dropout site ids) end to end through autograd
and compare it with the oracle.  The product never imports this file; on a GPU box the same orchestration runs on the real kernels
and `tests/test_train_gpu.py` checks those one by one.

`ROUND_BF16 = False` keeps everything in fp32 (exactness of the orchestration); True rounds where the kernels round (bf16 GEMM operands
and bf16 activations), which gives a CPU estimate of the numerical noise of the real path.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from oracle import philox

D = 768
ROUND_BF16 = False
REAL_DTYPES = False     # True: bf16 results are real torch.bfloat16 tensors (needed when the stand-ins feed the emulated kernels of tests/emu)


def _r(x: torch.Tensor) -> torch.Tensor:
    """bf16 storage: the real kernels hand bf16 tensors around; here they stay fp32 holding bf16-representable values."""
    if REAL_DTYPES:
        return x.to(torch.bfloat16)
    return x.to(torch.bfloat16).float() if ROUND_BF16 else x


def require_cuda(t, name):
    return None


def cast_bf16(x):
    return _r(x.float())


def gemm(a, w, bias, out=None, *, gelu=False, residual=None, out_f32=False, impl=None):
    y = a.float() @ w.float().t()
    if bias is not None:
        y = y + bias
    if gelu:
        y = F.gelu(y)
    if residual is not None:
        y = y + residual.reshape(-1, y.shape[1])
    y = y if out_f32 else _r(y)
    if out is not None:
        out.copy_(y)
        return out
    return y


def layernorm(x, gamma, beta, eps, out=None, *, rows=None, group=None, group_stride=None, offset=0, gamma2=None, beta2=None, eps2=0.0,
              out_f32=False):
    rows = x.shape[0] if rows is None else rows
    if group is None:
        group, group_stride = rows, rows
    r = torch.arange(rows)
    src = (r // group) * group_stride + offset + r % group
    y = F.layer_norm(x[src], (D,), gamma, beta, eps)
    if gamma2 is not None:
        y = F.layer_norm(y, (D,), gamma2, beta2, eps2)
    y = y if out_f32 else _r(y)
    if out is not None:
        out[:rows].copy_(y)
        return out
    return y


def sync_tokens(v, a, vw, vb, aw, ab, eps, off_tok, mod_tok, pos_emb, B, S):
    v = F.layer_norm(v.reshape(B, 8 * S, D), (D,), vw, vb, eps)
    a = F.layer_norm(a.reshape(B, 6 * S, D), (D,), aw, ab, eps)
    x = torch.cat([off_tok.reshape(1, 1, D).expand(B, 1, D), v, mod_tok.reshape(1, 1, D).expand(B, 1, D), a], dim=1) + pos_emb.reshape(1, -1, D)
    return x.reshape(-1, D).contiguous()


def sync_head(x, T, ln_w, ln_b, eps, W, b, B):
    return F.linear(F.layer_norm(x.reshape(B, T, D)[:, 0], (D,), ln_w, ln_b, eps), W, b)


def _mult(shape, p, seed, site):
    return torch.from_numpy(philox.dropout_multiplier(tuple(shape), p, seed, site))


def dropout(x, p, seed, site, *, residual=None, out=None, out_bf16=False):
    y = x * _mult(x.shape, p, seed, site)
    if residual is not None:
        y = y + residual
    y = _r(y) if out_bf16 else y
    if out is not None:
        out.copy_(y)
        return out
    return y


def gelu_fwd(x):
    return _r(F.gelu(x.float()))


def gelu_bwd(dy, x):
    x = x.float()
    cdf = 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))
    pdf = torch.exp(-0.5 * x * x) / math.sqrt(2.0 * math.pi)
    return _r(dy.float() * (cdf + x * pdf))


def transpose_bf16(x):
    R, C = x.shape
    out = torch.zeros((C, (R + 7) // 8 * 8), dtype=x.dtype)
    out[:, :R] = x.t()
    return out


def colsum(x):
    return x.float().sum(0)


def layernorm_bwd(dy, x, gamma, eps, *, dx=None, accumulate=False, rows=None, group=None, group_stride=None, offset=0):
    rows = x.shape[0] if rows is None else rows
    if group is None:
        group, group_stride = rows, rows
    r = torch.arange(rows)
    g = dy[(r // group) * group_stride + offset + r % group]
    xr = x[:rows]
    mean = xr.mean(-1, keepdim=True)
    rstd = torch.rsqrt(xr.var(-1, unbiased=False, keepdim=True) + eps)
    xhat = (xr - mean) * rstd
    gg = g * gamma
    res = rstd * (gg - gg.mean(-1, keepdim=True) - xhat * (gg * xhat).mean(-1, keepdim=True))
    if dx is None:
        dx = res
    elif accumulate:
        dx[:rows] += res
    else:
        dx[:rows] = res
    return dx, (g * xhat).sum(0), g.sum(0)


def _split(qkv, B, T, h, d):
    q, k, v = qkv.float().reshape(B, T, 3, h, d).permute(2, 0, 3, 1, 4)       # each (B, h, T, d)
    return q, k, v


def attention_train_fwd(qkv, B, T, n_heads, head_dim, scale, p, seed, site):
    q, k, v = _split(qkv, B, T, n_heads, head_dim)
    s2 = (q @ k.transpose(-1, -2)) * (scale * math.log2(math.e))
    mx = s2.max(-1, keepdim=True).values
    e = torch.exp2(s2 - mx)
    ssum = e.sum(-1, keepdim=True)
    pd = e / ssum * _mult((B, n_heads, T, T), p, seed, site)
    out = (pd @ v).permute(0, 2, 1, 3).reshape(B * T, n_heads * head_dim)
    return _r(out), (mx + torch.log2(ssum)).squeeze(-1).contiguous()


def attention_train_bwd(qkv, out, d_out, lse, B, T, n_heads, head_dim, scale, p, seed, site):
    q, k, v = _split(qkv, B, T, n_heads, head_dim)
    o = out.float().reshape(B, T, n_heads, head_dim).permute(0, 2, 1, 3)
    do = d_out.float().reshape(B, T, n_heads, head_dim).permute(0, 2, 1, 3)
    m = _mult((B, n_heads, T, T), p, seed, site)
    P = torch.exp2((q @ k.transpose(-1, -2)) * (scale * math.log2(math.e)) - lse.reshape(B, n_heads, T, 1))
    delta = (do * o).sum(-1, keepdim=True)
    dpd = do @ v.transpose(-1, -2)
    ds = P * (dpd * m - delta)
    dq = scale * (ds @ k)
    dk = scale * (ds.transpose(-1, -2) @ q)
    dv = (P * m).transpose(-1, -2) @ do
    dqkv = torch.stack([dq, dk, dv], dim=0).permute(1, 3, 0, 2, 4).reshape(B * T, 3 * n_heads * head_dim)
    return _r(dqkv)


def sync_head_bwd(x, T, ln_w, ln_b, eps, W, dlogits, B):
    row = x.reshape(B, T, D)[:, 0]
    mean = row.mean(-1, keepdim=True)
    rstd = torch.rsqrt(row.var(-1, unbiased=False, keepdim=True) + eps)
    xhat = (row - mean) * rstd
    y = xhat * ln_w + ln_b
    dyn = dlogits @ W
    gg = dyn * ln_w
    dx = torch.zeros_like(x).reshape(B, T, D)
    dx[:, 0] = rstd * (gg - gg.mean(-1, keepdim=True) - xhat * (gg * xhat).mean(-1, keepdim=True))
    return dx.reshape(B * T, D), (dyn * xhat).sum(0), dyn.sum(0), dlogits.t() @ y, dlogits.sum(0)


ALL = ('require_cuda', 'cast_bf16', 'gemm', 'layernorm', 'sync_tokens', 'sync_head', 'dropout', 'gelu_fwd', 'gelu_bwd',
       'transpose_bf16', 'colsum', 'layernorm_bwd', 'attention_train_fwd', 'attention_train_bwd', 'sync_head_bwd')


def install(monkeypatch, round_bf16: bool = False, names=ALL, real_dtypes: bool = False):
    """Route `ops.<kernel>` for the given names (default: everything the sync-module training path uses) to the CPU stand-ins above."""
    import sys
    from synchformer_b200 import ops
    me = sys.modules[__name__]
    monkeypatch.setattr(me, 'ROUND_BF16', round_bf16)
    monkeypatch.setattr(me, 'REAL_DTYPES', real_dtypes)
    for name in names:
        monkeypatch.setattr(ops, name, getattr(me, name))
