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


def gemm(a, w, bias, out=None, *, gelu=False, residual=None, out_f32=False, impl=None, ln_fold=None, emit_ln=None):
    """restates the sfb_gemm_bf16 / sfb_gemm_bf16_ln contract (include/synchformer_b200.h)"""
    y = a.float() @ w.float().t()
    if ln_fold is not None:                      # rstd (acc - mean colsum) + bias from the partial (sum, sum of squares) of the A rows
        stats, colsum, eps = ln_fold
        K = a.shape[1]
        s1, s2 = stats[:a.shape[0], :, 0].sum(1, keepdim=True), stats[:a.shape[0], :, 1].sum(1, keepdim=True)
        mean = s1 / K
        rstd = torch.rsqrt((s2 / K - mean * mean).clamp_min(0.0) + eps)
        y = (y - mean * colsum.unsqueeze(0)) * rstd
    if bias is not None:
        y = y + bias
    if gelu:
        y = F.gelu(y)
    if residual is not None:
        y = y + residual.reshape(-1, y.shape[1])
    if emit_ln is not None:
        xb, stats_out = emit_ln
        xb.copy_(_r(y))
        g = y.reshape(y.shape[0], y.shape[1] // 64, 64)
        stats_out.copy_(torch.stack([g.sum(-1), (g * g).sum(-1)], dim=-1))
    y = y if out_f32 else _r(y)
    if out is not None:
        out.copy_(y)
        return out
    return y


def rowstats_cast(x, xb=None, stats=None):
    x = x.float()
    if xb is None:
        xb = empty_bf16(x.shape, x.device)
    if stats is None:
        stats = torch.empty((x.shape[0], 1, 2))
    xb.copy_(_r(x))
    stats.copy_(torch.stack([x.sum(-1, keepdim=True), (x * x).sum(-1, keepdim=True)], dim=-1))
    return xb, stats


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


def transpose_bf16(x, pad_to=8):
    R, C = x.shape
    out = torch.zeros((C, (R + pad_to - 1) // pad_to * pad_to), dtype=x.dtype)
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


# ---- encoder forward kernels (needed by the N1 tests; verified on hardware in round 1, restated here from their contracts) ----------
def im2col_video(vis, out=None):
    n = vis.shape[0]
    x = vis.float().reshape(n, 8, 2, 3, 14, 16, 14, 16).permute(0, 1, 4, 6, 3, 2, 5, 7).reshape(n * 1568, 1536)   # K order (c, dt, dy, dx)
    return _r(x.contiguous())


def video_tokens(patch, cls_token, pos_embed, temp_embed, n, out=None):
    pos, tmp = pos_embed.reshape(197, D), temp_embed.reshape(8, D)
    tok = patch.reshape(n, 8, 196, D) + pos[1:].unsqueeze(0).unsqueeze(0) + tmp.reshape(1, 8, 1, D)
    cls = (cls_token.reshape(1, 1, D) + pos[0].reshape(1, 1, D)).expand(n, 1, D)
    return torch.cat([cls, tok.reshape(n, 1568, D)], dim=1).reshape(n * 1569, D).contiguous()


def im2col_ast(spec):
    n = spec.shape[0]
    return _r(spec.unfold(1, 16, 10).unfold(2, 16, 10).reshape(n * 72, 256).contiguous())


def ast_tokens(patch, cls_token, dist_token, pos_embed, n):
    x = torch.cat([cls_token.reshape(1, 1, D).expand(n, 1, D), dist_token.reshape(1, 1, D).expand(n, 1, D), patch.reshape(n, 72, D)], dim=1)
    return (x + pos_embed.reshape(1, 74, D)).reshape(n * 74, D).contiguous()


def attention(q, k, v, out, *, q_strides, kv_strides, o_strides, n_outer, n_inner, n_heads, head_dim, Lq, Lk, scale, k_prefix=None, v_prefix=None,
              prefix_outer=0, impl=None, q_extra=None, q_extra_outer=0, extra_out=None, extra_out_outer=0):
    """sfb_attention on strided views (see sfb_attn_desc); the optional fused extra query is reported as not fused (the caller then issues
    it as its own call, which is the documented fallback of ops.attention)."""
    hd = head_dim
    view = lambda t, st, L: t.as_strided((n_outer, n_inner, n_heads, L, hd), (st[0], st[1], hd, st[2], 1))
    qq, kk, vv = view(q, q_strides, Lq).float(), view(k, kv_strides, Lk).float(), view(v, kv_strides, Lk).float()
    if k_prefix is not None:
        pk = k_prefix.as_strided((n_outer, 1, n_heads, 1, hd), (prefix_outer, 0, hd, 0, 1)).float().expand(n_outer, n_inner, n_heads, 1, hd)
        pv = v_prefix.as_strided((n_outer, 1, n_heads, 1, hd), (prefix_outer, 0, hd, 0, 1)).float().expand(n_outer, n_inner, n_heads, 1, hd)
        kk, vv = torch.cat([pk, kk], dim=3), torch.cat([pv, vv], dim=3)
    o = torch.softmax(qq @ kk.transpose(-1, -2) * scale, dim=-1) @ vv
    view(out, o_strides, Lq).copy_(o)
    return False


# ---- N1 backward kernels --------------------------------------------------------------------------------------------------------
def _views(n_outer, n_inner, n_heads, hd):
    return lambda t, st, L: t.as_strided((n_outer, n_inner, n_heads, L, hd), (st[0], st[1], hd, st[2], 1))


def attention_bwd(q, k, v, out, d_out, dq, dk, dv, *, q_strides, kv_strides, o_strides, n_outer, n_inner, n_heads, head_dim, Lq, Lk, scale,
                  k_prefix=None, v_prefix=None, prefix_outer=0, impl=0):
    """sfb_attention_bwd restated with autograd on the dense definition; prefix gradients are returned per problem."""
    hd = head_dim
    view = _views(n_outer, n_inner, n_heads, hd)
    with torch.enable_grad():
        qq = view(q, q_strides, Lq).float().clone().requires_grad_(True)
        kk = view(k, kv_strides, Lk).float().clone().requires_grad_(True)
        vv = view(v, kv_strides, Lk).float().clone().requires_grad_(True)
        leaves, K, V = [qq, kk, vv], kk, vv
        if k_prefix is not None:
            pre = lambda t: t.as_strided((n_outer, 1, n_heads, 1, hd), (prefix_outer, 0, hd, 0, 1)).float().expand(n_outer, n_inner, n_heads, 1, hd)
            pk, pv = pre(k_prefix).clone().requires_grad_(True), pre(v_prefix).clone().requires_grad_(True)
            leaves += [pk, pv]
            K, V = torch.cat([pk, kk], dim=3), torch.cat([pv, vv], dim=3)
        o = torch.softmax(qq @ K.transpose(-1, -2) * scale, dim=-1) @ V
        grads = torch.autograd.grad(o, leaves, view(d_out, o_strides, Lq).float())
    view(dq, q_strides, Lq).copy_(grads[0])
    view(dk, kv_strides, Lk).copy_(grads[1])
    view(dv, kv_strides, Lk).copy_(grads[2])
    if k_prefix is None:
        return None
    return torch.stack([grads[3][:, :, :, 0], grads[4][:, :, :, 0]], dim=3).permute(1, 0, 2, 3, 4).contiguous()      # (inner, outer, heads, 2, hd)


def attention_bwd_global_query(q, k, v, out, d_out, dq, dk, dv, *, q_outer, kv_outer, kv_row, o_outer, n_outer, n_heads, head_dim, Lk, scale,
                               prefix_grad=None):
    hd = head_dim
    rowv = lambda t, st: t.as_strided((n_outer, n_heads, 1, hd), (st, hd, 0, 1))
    kvv = lambda t: t.as_strided((n_outer, n_heads, Lk, hd), (kv_outer, hd, kv_row, 1))
    with torch.enable_grad():
        qq = rowv(q, q_outer).float().clone().requires_grad_(True)
        kk, vv = kvv(k).float().clone().requires_grad_(True), kvv(v).float().clone().requires_grad_(True)
        o = torch.softmax(qq @ kk.transpose(-1, -2) * scale, dim=-1) @ vv
        gq, gk, gv = torch.autograd.grad(o, [qq, kk, vv], rowv(d_out, o_outer).float())
    if prefix_grad is not None:
        pg = prefix_grad.reshape(n_outer, n_heads, 2, hd)
        gk[:, :, 0] += pg[:, :, 0]
        gv[:, :, 0] += pg[:, :, 1]
    rowv(dq, q_outer).copy_(gq)
    kvv(dk).copy_(kvv(dk).float() + gk)
    kvv(dv).copy_(kvv(dv).float() + gv)


def droppath(x, rows_per_sample, p, seed, site, *, residual=None, out=None, out_bf16=False):
    n_samples = x.shape[0] // rows_per_sample
    m = _mult((n_samples,), p, seed, site).repeat_interleave(rows_per_sample).unsqueeze(1)
    y = x * m
    if residual is not None:
        y = y + residual
    y = _r(y) if out_bf16 else y
    if out is not None:
        out.copy_(y)
        return out
    return y


def gather_rows_bf16(x, rows, group=None, group_stride=None, offset=0):
    if group is None:
        group, group_stride = rows, rows
    r = torch.arange(rows)
    return _r(x[(r // group) * group_stride + offset + r % group].contiguous())


def empty_bf16(shape, device):
    return torch.empty(shape, device=device, dtype=torch.bfloat16 if (REAL_DTYPES or ROUND_BF16) else torch.float32)


def mean_tokens(x):
    return x.float().mean(dim=1)


def mean_tokens_bwd(dout, T):
    return (dout / T).unsqueeze(1).expand(-1, T, -1).contiguous()


def l2_normalize(x):
    inv = 1.0 / x.norm(dim=-1).clamp_min(1e-12)
    return x * inv.unsqueeze(1), inv


def l2_normalize_bwd(xn, inv_norm, dxn):
    return (dxn - xn * (xn * dxn).sum(-1, keepdim=True)) * inv_norm.unsqueeze(1)


def contrastive_loss(vn, an, vn_all, an_all, scale):
    n, N = vn.shape[0], vn_all.shape[0]
    sims = torch.stack([vn @ an_all.mT, an @ vn_all.mT]) / scale                 # (2, n, N)
    eye = torch.eye(n, N)
    G = (torch.softmax(sims, dim=-1) - eye) * (0.5 / n)
    loss = (-(torch.log_softmax(sims, dim=-1) * eye).sum(-1)).sum() * (0.5 / n)
    dscale = -(G * sims).sum() / scale
    return loss.reshape(1), dscale.reshape(1), G.contiguous()


def contrastive_loss_bwd(vn, an, vn_all, an_all, G, scale, upstream, dscale):
    local = vn_all is None
    kv, ka = (vn, an) if local else (vn_all, an_all)
    coef = upstream / scale
    d_vn, d_an = G[0] @ ka * coef, G[1] @ kv * coef
    d_ka, d_kv = G[0].mT @ vn * coef, G[1].mT @ an * coef
    if local:
        return d_vn + d_kv, d_an + d_ka, None, None, dscale * upstream
    return d_vn, d_an, d_kv, d_ka, dscale * upstream


CONTRASTIVE = ('mean_tokens', 'mean_tokens_bwd', 'l2_normalize', 'l2_normalize_bwd', 'contrastive_loss', 'contrastive_loss_bwd')
N1_BWD = ('attention_bwd', 'attention_bwd_global_query', 'droppath', 'gather_rows_bf16', 'empty_bf16')
ENCODER_FWD = ('im2col_video', 'video_tokens', 'im2col_ast', 'ast_tokens', 'attention', 'rowstats_cast')

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
