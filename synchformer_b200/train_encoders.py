"""backward of the two feature extractors (stage-I contrastive training of `AVCLIP`,
model/modules/feat_extractors/train_clip_src/open_clip/model.py:474-527 driven by training/train.py:122-154; also what stage II needs
when `is_trainable: True`).

`motionformer_features(module, vis)` and `ast_features(module, spec)` are the differentiable versions of `MotionFormer.encode` /
`AST.encode`: a `torch.autograd.Function` each, whose forward runs the same kernels as inference while keeping the per-layer
activations, and whose backward is written by hand on the kernels of csrc/train.cu, attention_train.cu, attention_bwd.cu and the tcgen05
GEMM (dX = dY W, dW = dY^T X on transposed operands).  Only parameter gradients are produced (the inputs are pixels / spectrograms).

Reference semantics kept: DropPath (stochastic depth, rate 0.2 * i / 11 in block i; timm `DropPath` at vit_helper.py:356,371,375) on
the space-attention and MLP branches of the Motionformer in train mode - one keep decision per segment from the counter-based generator
of csrc/philox.cuh, site 2 i (space) / 2 i + 1 (MLP); every other dropout of both towers has rate 0 (divided_224_16x4.yaml:57-61,
ASTConfig defaults).  The CLS aggregators are evaluated for the CLS row only, as in inference (exact), and so is their backward.

Numerics as the reference under autocast(bf16): bf16 GEMM operands / activations, fp32 accumulation, LayerNorm, softmax and residual
stream; fp32 gradients.  PyTorch holds memory and the autograd graph; the only torch arithmetic is on bias-sized vectors
(sums of two partial gradients of the same parameter) and the 1/8, 1/6 of the time-average pooling.
"""
import math
from typing import Dict, List, Optional, Tuple

import torch

from . import ops, train
from .train import _t

D = 768
EPS_V, EPS_A = 1e-6, 1e-12
V_TOK, V_SPACE, V_FRAMES = 1569, 196, 8
A_TOK, A_F, A_T = 74, 12, 6
DROP_PATH_RATE = 0.2          # divided_224_16x4.yaml:59, linspace over the 12 blocks (video_model_builder.py:86-87)


def drop_path_rate(i: int, depth: int = 12) -> float:
    return DROP_PATH_RATE * i / (depth - 1)


def _lin_bwd(dy_b: torch.Tensor, x_b: torch.Tensor, w_b: Optional[torch.Tensor], dx_f32: bool):
    """y = x W^T + b.  dy_b (M, N) bf16, x_b (M, K) bf16, w_b (N, K) bf16 -> (dW (N, K) fp32, db (N,) fp32, dx (M, K) bf16 | fp32 | None)."""
    dw = ops.gemm(_t(dy_b), _t(x_b), None, out_f32=True)
    db = ops.colsum(dy_b)
    dx = None if w_b is None else ops.gemm(dy_b, _t(w_b), None, out_f32=dx_f32)
    return dw, db, dx


# ------------------------------------------------------------------------------------------------------------------
# CLS aggregator (BaseEncoderLayer, motionformer.py:301-334), CLS row only
# ------------------------------------------------------------------------------------------------------------------
def _agg_forward(P, W, prefix: str, kv_src: torch.Tensor, n_outer: int, n_inner: int, kv_strides_rows: Tuple[int, int, int], Lk: int):
    """kv_src (rows, 768) bf16 = norm1(final_norm(tokens)).  Returns (out (n_outer * n_inner, 768) fp32, saved)."""
    G = n_outer * n_inner
    dev = kv_src.device
    kv = ops.gemm(kv_src, W[prefix + 'kv_w'], P[prefix + 'self_attn.in_proj_bias'][D:])                                  # (rows, 1536)
    cls_ln = ops.layernorm(P[prefix + 'cls_token'].view(1, D), P[prefix + 'norm1.weight'], P[prefix + 'norm1.bias'], EPS_V)
    cls_qkv = ops.gemm(cls_ln, W[prefix + 'in_w'], P[prefix + 'self_attn.in_proj_bias'])                                 # (1, 2304)
    ao = ops.empty_bf16((G, D), dev)
    kvs = tuple(s * 2 * D for s in kv_strides_rows)
    ops.attention(cls_qkv, kv, kv[:, D:], ao, q_strides=(0, 0, 0), kv_strides=kvs, o_strides=(n_inner * D, D, D), n_outer=n_outer,
                  n_inner=n_inner, n_heads=12, head_dim=64, Lq=1, Lk=Lk, scale=0.125, k_prefix=cls_qkv[:, D:], v_prefix=cls_qkv[:, 2 * D:],
                  prefix_outer=0)
    y0 = ops.gemm(ao, W[prefix + 'out_w'], P[prefix + 'self_attn.out_proj.bias'], residual=P[prefix + 'cls_token'].view(1, D), out_f32=True)
    h_ln = ops.layernorm(y0, P[prefix + 'norm2.weight'], P[prefix + 'norm2.bias'], EPS_V)
    hpre = ops.gemm(h_ln, W[prefix + 'l1_w'], P[prefix + 'linear1.bias'])
    hid = ops.gelu_fwd(hpre)
    out = ops.gemm(hid, W[prefix + 'l2_w'], P[prefix + 'linear2.bias'], residual=y0, out_f32=True)
    return out, dict(kv_src=kv_src, kv=kv, cls_ln=cls_ln, cls_qkv=cls_qkv, ao=ao, y0=y0, h_ln=h_ln, hpre=hpre, hid=hid, n_outer=n_outer,
                     n_inner=n_inner, kvs=kvs, Lk=Lk)


def _agg_backward(P, W, prefix: str, sv, d_out: torch.Tensor, y1: torch.Tensor, g: Dict[str, torch.Tensor]) -> torch.Tensor:
    """d_out (G, 768) fp32 -> gradient w.r.t. y1 (rows, 768) fp32 (the aggregator's input tokens, before its norm1); parameter
    gradients go to g[prefix + name]."""
    n_outer, n_inner, Lk = sv['n_outer'], sv['n_inner'], sv['Lk']
    G = n_outer * n_inner
    dev = d_out.device
    # out = linear2(gelu(linear1(norm2(y0)))) + y0
    dy_b = ops.cast_bf16(d_out)
    g[prefix + 'linear2.weight'], g[prefix + 'linear2.bias'], dhid = _lin_bwd(dy_b, sv['hid'], W[prefix + 'l2_w'], False)
    dpre = ops.gelu_bwd(dhid, sv['hpre'])
    g[prefix + 'linear1.weight'], g[prefix + 'linear1.bias'], dh_ln = _lin_bwd(dpre, sv['h_ln'], W[prefix + 'l1_w'], True)
    dy0 = d_out.clone()
    _, g[prefix + 'norm2.weight'], g[prefix + 'norm2.bias'] = ops.layernorm_bwd(dh_ln, sv['y0'], P[prefix + 'norm2.weight'], EPS_V, dx=dy0,
                                                                                 accumulate=True)
    # y0 = out_proj(ao) + cls_token
    dy0_b = ops.cast_bf16(dy0)
    g[prefix + 'self_attn.out_proj.weight'], g[prefix + 'self_attn.out_proj.bias'], dao = _lin_bwd(dy0_b, sv['ao'], W[prefix + 'out_w'], False)
    d_cls_res = ops.colsum(dy0)                                                   # residual branch of the learned CLS token
    # attention of the single CLS query of every group; q is one shared row, so dq comes out per group and is summed afterwards
    cls_qkv, kv = sv['cls_qkv'], sv['kv']
    q_rep = cls_qkv[:, :D].expand(G, D).contiguous()
    dq_rep = ops.empty_bf16((G, D), dev)
    dkv = torch.empty_like(kv)
    part = ops.attention_bwd(q_rep, kv, kv[:, D:], sv['ao'], dao, dq_rep, dkv, dkv[:, D:], q_strides=(n_inner * D, D, D), kv_strides=sv['kvs'],
                             o_strides=(n_inner * D, D, D), n_outer=n_outer, n_inner=n_inner, n_heads=12, head_dim=64, Lq=1, Lk=Lk, scale=0.125,
                             k_prefix=cls_qkv[:, D:], v_prefix=cls_qkv[:, 2 * D:], prefix_outer=0)
    d_cls_kv = ops.colsum(part.view(G, 12 * 2 * 64)).view(12, 2, 64)              # prefix row shared by ALL groups
    d_cls_qkv = torch.cat([ops.colsum(dq_rep), d_cls_kv[:, 0].reshape(-1), d_cls_kv[:, 1].reshape(-1)]).view(1, 3 * D)
    # kv = in_proj[768:](kv_src);  cls_qkv = in_proj(norm1(cls_token))
    g_in_w = torch.zeros((3 * D, D), device=dev, dtype=torch.float32)
    g_kv_w, g_kv_b, dkv_src = _lin_bwd(dkv, sv['kv_src'], W[prefix + 'kv_w'], True)
    g_in_w[D:].copy_(g_kv_w)
    d_cls_qkv_b = ops.cast_bf16(d_cls_qkv)
    ops.gemm(_t(d_cls_qkv_b), _t(sv['cls_ln']), None, out=g_in_w, residual=g_in_w, out_f32=True)     # += d_cls_qkv^T cls_ln  (K padded 1 -> 8)
    g[prefix + 'self_attn.in_proj_weight'] = g_in_w
    g_in_b = d_cls_qkv.view(-1).clone()
    g_in_b[D:] += g_kv_b
    g[prefix + 'self_attn.in_proj_bias'] = g_in_b
    d_cls_ln = ops.gemm(d_cls_qkv_b, _t(W[prefix + 'in_w']), None, out_f32=True)                      # (1, 768)
    # norm1 is applied to the tokens (kv_src = norm1(y1)) and to the CLS token
    d_cls_tok, dg_a, db_a = ops.layernorm_bwd(d_cls_ln, P[prefix + 'cls_token'].view(1, D), P[prefix + 'norm1.weight'], EPS_V)
    dy1, dg_b, db_b = ops.layernorm_bwd(dkv_src, y1, P[prefix + 'norm1.weight'], EPS_V)
    g[prefix + 'norm1.weight'], g[prefix + 'norm1.bias'] = dg_a + dg_b, db_a + db_b
    g[prefix + 'cls_token'] = (d_cls_res + d_cls_tok.view(-1)).view(1, 1, D)
    return dy1


# ------------------------------------------------------------------------------------------------------------------
# pre-norm ViT block pieces shared by both towers
# ------------------------------------------------------------------------------------------------------------------
def _mlp_backward(P, W, dx: torch.Tensor, dy_b: torch.Tensor, x_in: torch.Tensor, ln: torch.Tensor, hpre: torch.Tensor, hid: torch.Tensor,
                  names: Tuple[str, str, str], wkeys: Tuple[str, str], eps: float, g: Dict[str, torch.Tensor]):
    """x_out = x_in + [drop] fc2(gelu(fc1(norm(x_in)))); dy_b = bf16 gradient of the branch output; accumulates into dx."""
    fc1, fc2, norm = names
    g[fc2 + '.weight'], g[fc2 + '.bias'], dhid = _lin_bwd(dy_b, hid, W[wkeys[1]], False)
    dpre = ops.gelu_bwd(dhid, hpre)
    g[fc1 + '.weight'], g[fc1 + '.bias'], dln = _lin_bwd(dpre, ln, W[wkeys[0]], True)
    _, g[norm + '.weight'], g[norm + '.bias'] = ops.layernorm_bwd(dln, x_in, P[norm + '.weight'], eps, dx=dx, accumulate=True)


# ------------------------------------------------------------------------------------------------------------------
# AST tower (ast.py:137-279, hf_src/modeling_ast.py:83-117, 145-184, 294-322, 543)
# ------------------------------------------------------------------------------------------------------------------
def _ast_forward(m, spec: torch.Tensor):
    """spec (n, 128, 66) fp32 -> (feats (n, 6, 768) fp32, saved)."""
    P, W = m.weights()
    n = spec.shape[0]
    e = 'ast.embeddings.'
    a = ops.im2col_ast(spec)
    patch = ops.gemm(a, W['pe_w'], P[e + 'patch_embeddings.projection.bias'], out_f32=True)
    x = ops.ast_tokens(patch, P[e + 'cls_token'], P[e + 'distillation_token'], P[e + 'position_embeddings'], n)          # (n*74, 768) fp32
    layers = []
    for i in range(12):                                                                                                   # ASTLayer modeling_ast.py:294-322
        l = f'ast.encoder.layer.{i}.'
        ln1 = ops.layernorm(x, P[l + 'layernorm_before.weight'], P[l + 'layernorm_before.bias'], EPS_A)
        qkv = ops.gemm(ln1, W[l + 'qkv'], W[l + 'qkv_b'])
        att, lse = ops.attention_train_fwd(qkv, n, A_TOK, 12, 64, 0.125, 0.0, 0, 0)
        x_mid = ops.gemm(att, W[l + 'o'], P[l + 'attention.output.dense.bias'], residual=x, out_f32=True)
        ln2 = ops.layernorm(x_mid, P[l + 'layernorm_after.weight'], P[l + 'layernorm_after.bias'], EPS_A)
        hpre = ops.gemm(ln2, W[l + 'fc1'], P[l + 'intermediate.dense.bias'])
        hid = ops.gelu_fwd(hpre)
        x_out = ops.gemm(hid, W[l + 'fc2'], P[l + 'output.dense.bias'], residual=x_mid, out_f32=True)
        layers.append((x, ln1, qkv, att, lse, x_mid, ln2, hpre, hid))
        x = x_out
    g = 'freq_attn_agg.'
    y1 = ops.layernorm(x, P['ast.layernorm.weight'], P['ast.layernorm.bias'], EPS_A, rows=n * 72, group=72, group_stride=A_TOK, offset=2, out_f32=True)
    kv_src = ops.layernorm(y1, P[g + 'norm1.weight'], P[g + 'norm1.bias'], EPS_V)
    feats, agg = _agg_forward(P, W, g, kv_src, n, A_T, (72, 1, A_T), A_F)                                                # (n*6, 768)
    return feats.view(n, A_T, D), dict(n=n, a=a, layers=layers, x_final=x, y1=y1, agg=agg)


def _ast_backward(m, sv, d_feats: torch.Tensor) -> Dict[str, torch.Tensor]:
    """d_feats (n, 6, 768) fp32 -> {parameter name: gradient}."""
    P, W = m.weights()
    n = sv['n']
    g: Dict[str, torch.Tensor] = {}
    dy1 = _agg_backward(P, W, 'freq_attn_agg.', sv['agg'], d_feats.reshape(n * A_T, D).contiguous(), sv['y1'], g)
    # final LayerNorm over the 72 patch tokens of each segment (cls / distillation rows get no gradient)
    x_sel = sv['x_final'].view(n, A_TOK, D)[:, 2:].reshape(n * 72, D)
    dx_sel, g['ast.layernorm.weight'], g['ast.layernorm.bias'] = ops.layernorm_bwd(dy1, x_sel, P['ast.layernorm.weight'], EPS_A)
    dx = torch.zeros((n * A_TOK, D), device=dy1.device, dtype=torch.float32)
    dx.view(n, A_TOK, D)[:, 2:].copy_(dx_sel.view(n, 72, D))
    for i in reversed(range(12)):
        l = f'ast.encoder.layer.{i}.'
        x_in, ln1, qkv, att, lse, x_mid, ln2, hpre, hid = sv['layers'][i]
        _mlp_backward(P, W, dx, ops.cast_bf16(dx), x_mid, ln2, hpre, hid, (l + 'intermediate.dense', l + 'output.dense', l + 'layernorm_after'),
                      (l + 'fc1', l + 'fc2'), EPS_A, g)
        o = l + 'attention.output.dense'
        g[o + '.weight'], g[o + '.bias'], datt = _lin_bwd(ops.cast_bf16(dx), att, W[l + 'o'], False)
        dqkv = ops.attention_train_bwd(qkv, att, datt, lse, n, A_TOK, 12, 64, 0.125, 0.0, 0, 0)
        dw, db, dln1 = _lin_bwd(dqkv, ln1, W[l + 'qkv'], True)
        for k, name in enumerate(('query', 'key', 'value')):
            g[l + f'attention.attention.{name}.weight'] = dw[k * D:(k + 1) * D]
            g[l + f'attention.attention.{name}.bias'] = db[k * D:(k + 1) * D]
        _, g[l + 'layernorm_before.weight'], g[l + 'layernorm_before.bias'] = ops.layernorm_bwd(dln1, x_in, P[l + 'layernorm_before.weight'], EPS_A,
                                                                                                  dx=dx, accumulate=True)
    # embeddings: x0 = [cls, dist, conv(patches)] + pos   (modeling_ast.py:83-93, 113-117)
    e = 'ast.embeddings.'
    dpos = ops.colsum(dx.view(n, A_TOK * D)).view(A_TOK, D)
    g[e + 'position_embeddings'] = dpos.view(1, A_TOK, D)
    g[e + 'cls_token'] = dpos[0].clone().view(1, 1, D)
    g[e + 'distillation_token'] = dpos[1].clone().view(1, 1, D)
    dpatch_b = ops.gather_rows_bf16(dx, n * 72, 72, A_TOK, 2)
    dw, g[e + 'patch_embeddings.projection.bias'], _ = _lin_bwd(dpatch_b, sv['a'], None, False)
    g[e + 'patch_embeddings.projection.weight'] = dw.view(D, 1, 16, 16)
    return g


# ------------------------------------------------------------------------------------------------------------------
# Motionformer tower (motionformer.py:182-272, vit_helper.py:100-158, 364-376, video_model_builder.py:174-274)
# ------------------------------------------------------------------------------------------------------------------
def _divided_attention_bwd(qkv: torch.Tensor, att: torch.Tensor, datt: torch.Tensor, n: int, mode: str) -> torch.Tensor:
    """Backward of MotionFormer._divided_attention on the fused (n*1569, 2304) layout -> dqkv (same layout, bf16)."""
    row, seg = 3 * D, V_TOK * 3 * D
    dqkv = torch.zeros_like(qkv)          # the CLS rows of dK / dV are only ever accumulated into
    q, k, v = qkv, qkv[:, D:], qkv[:, 2 * D:]
    dq, dk, dv = dqkv, dqkv[:, D:], dqkv[:, 2 * D:]
    if mode == 'time':
        kw = dict(q_strides=(seg, row, V_SPACE * row), kv_strides=(seg, row, V_SPACE * row), o_strides=(V_TOK * D, D, V_SPACE * D), n_inner=V_SPACE,
                  Lq=V_FRAMES, Lk=V_FRAMES)
    else:
        kw = dict(q_strides=(seg, V_SPACE * row, row), kv_strides=(seg, V_SPACE * row, row), o_strides=(V_TOK * D, V_SPACE * D, D), n_inner=V_FRAMES,
                  Lq=V_SPACE, Lk=V_SPACE)
    part = ops.attention_bwd(q[1:], k[1:], v[1:], att[1:], datt[1:], dq[1:], dk[1:], dv[1:], n_outer=n, n_heads=12, head_dim=64, scale=0.125,
                             k_prefix=k, v_prefix=v, prefix_outer=seg, **kw)
    prefix_grad = ops.colsum(part.view(kw['n_inner'], n * 12 * 2 * 64))            # CLS key / value: summed over the inner problems
    # the CLS query attends to all 1569 keys (vit_helper.py:124): adds its share to every dK / dV row and writes dq of token 0
    ops.attention_bwd_global_query(q, k, v, att, datt, dq, dk, dv, q_outer=seg, kv_outer=seg, kv_row=row, o_outer=V_TOK * D, n_outer=n, n_heads=12,
                                   head_dim=64, Lk=V_TOK, scale=0.125, prefix_grad=prefix_grad)
    return dqkv


def _motionformer_forward(m, vis: torch.Tensor, seed: int, stochastic: bool):
    """vis (n, 16, 3, 224, 224) -> (feats (n, 8, 768) fp32, saved)."""
    P, W = m.weights()
    a = ops.im2col_video(vis)
    n = a.shape[0] // 1568
    patch = ops.gemm(a, W['pe_w'], P['patch_embed_3d.proj.bias'], out_f32=True)
    x = ops.video_tokens(patch, P['cls_token'], P['pos_embed'], P['temp_embed'], n)                                       # (n*1569, 768) fp32
    M = n * V_TOK
    blocks = []
    for i in range(12):                                                                                                   # vit_helper.py:364-376
        b = f'blocks.{i}.'
        p = drop_path_rate(i) if stochastic else 0.0
        ln3 = ops.layernorm(x, P[b + 'norm3.weight'], P[b + 'norm3.bias'], EPS_V)
        qkv_t = ops.gemm(ln3, W[b + 'timeattn.qkv'], P[b + 'timeattn.qkv.bias'])
        att_t = ops.empty_bf16((M, D), x.device)
        m._divided_attention(qkv_t, att_t, n, 'time')
        x1 = ops.gemm(att_t, W[b + 'timeattn.proj'], P[b + 'timeattn.proj.bias'], residual=x, out_f32=True)
        ln1 = ops.layernorm(x1, P[b + 'norm1.weight'], P[b + 'norm1.bias'], EPS_V)
        qkv_s = ops.gemm(ln1, W[b + 'attn.qkv'], P[b + 'attn.qkv.bias'])
        att_s = ops.empty_bf16((M, D), x.device)
        m._divided_attention(qkv_s, att_s, n, 'space')
        if p > 0:
            y = ops.gemm(att_s, W[b + 'attn.proj'], P[b + 'attn.proj.bias'], out_f32=True)
            x2 = ops.droppath(y, V_TOK, p, seed, 2 * i, residual=x1)
        else:
            x2 = ops.gemm(att_s, W[b + 'attn.proj'], P[b + 'attn.proj.bias'], residual=x1, out_f32=True)
        ln2 = ops.layernorm(x2, P[b + 'norm2.weight'], P[b + 'norm2.bias'], EPS_V)
        hpre = ops.gemm(ln2, W[b + 'mlp.fc1'], P[b + 'mlp.fc1.bias'])
        hid = ops.gelu_fwd(hpre)
        if p > 0:
            y = ops.gemm(hid, W[b + 'mlp.fc2'], P[b + 'mlp.fc2.bias'], out_f32=True)
            x3 = ops.droppath(y, V_TOK, p, seed, 2 * i + 1, residual=x2)
        else:
            x3 = ops.gemm(hid, W[b + 'mlp.fc2'], P[b + 'mlp.fc2.bias'], residual=x2, out_f32=True)
        blocks.append((x, ln3, qkv_t, att_t, x1, ln1, qkv_s, att_s, x2, ln2, hpre, hid, p))
        x = x3
    g = 'spatial_attn_agg.'
    y1 = ops.layernorm(x, P['norm.weight'], P['norm.bias'], EPS_V, rows=n * 1568, group=1568, group_stride=V_TOK, offset=1, out_f32=True)
    kv_src = ops.layernorm(y1, P[g + 'norm1.weight'], P[g + 'norm1.bias'], EPS_V)
    feats, agg = _agg_forward(P, W, g, kv_src, n, V_FRAMES, (1568, V_SPACE, 1), V_SPACE)                                  # (n*8, 768)
    return feats.view(n, V_FRAMES, D), dict(n=n, a=a, blocks=blocks, x_final=x, y1=y1, agg=agg, seed=seed)


def _motionformer_backward(m, sv, d_feats: torch.Tensor) -> Dict[str, torch.Tensor]:
    P, W = m.weights()
    n, seed = sv['n'], sv['seed']
    M = n * V_TOK
    g: Dict[str, torch.Tensor] = {}
    dy1 = _agg_backward(P, W, 'spatial_attn_agg.', sv['agg'], d_feats.reshape(n * V_FRAMES, D).contiguous(), sv['y1'], g)
    x_sel = sv['x_final'].view(n, V_TOK, D)[:, 1:].reshape(n * 1568, D)
    dx_sel, g['norm.weight'], g['norm.bias'] = ops.layernorm_bwd(dy1, x_sel, P['norm.weight'], EPS_V)
    dx = torch.zeros((M, D), device=dy1.device, dtype=torch.float32)
    dx.view(n, V_TOK, D)[:, 1:].copy_(dx_sel.view(n, 1568, D))
    for i in reversed(range(12)):
        b = f'blocks.{i}.'
        x_in, ln3, qkv_t, att_t, x1, ln1, qkv_s, att_s, x2, ln2, hpre, hid, p = sv['blocks'][i]
        # x3 = x2 + drop_path(mlp(norm2(x2)))
        dy_b = ops.droppath(dx, V_TOK, p, seed, 2 * i + 1, out_bf16=True) if p > 0 else ops.cast_bf16(dx)
        _mlp_backward(P, W, dx, dy_b, x2, ln2, hpre, hid, (b + 'mlp.fc1', b + 'mlp.fc2', b + 'norm2'), (b + 'mlp.fc1', b + 'mlp.fc2'), EPS_V, g)
        # x2 = x1 + drop_path(attn(norm1(x1)))        (space)
        dy_b = ops.droppath(dx, V_TOK, p, seed, 2 * i, out_bf16=True) if p > 0 else ops.cast_bf16(dx)
        g[b + 'attn.proj.weight'], g[b + 'attn.proj.bias'], datt = _lin_bwd(dy_b, att_s, W[b + 'attn.proj'], False)
        dqkv = _divided_attention_bwd(qkv_s, att_s, datt, n, 'space')
        g[b + 'attn.qkv.weight'], g[b + 'attn.qkv.bias'], dln = _lin_bwd(dqkv, ln1, W[b + 'attn.qkv'], True)
        _, g[b + 'norm1.weight'], g[b + 'norm1.bias'] = ops.layernorm_bwd(dln, x1, P[b + 'norm1.weight'], EPS_V, dx=dx, accumulate=True)
        # x1 = x + timeattn(norm3(x))                 (time; no drop path, vit_helper.py:366-368)
        g[b + 'timeattn.proj.weight'], g[b + 'timeattn.proj.bias'], datt = _lin_bwd(ops.cast_bf16(dx), att_t, W[b + 'timeattn.proj'], False)
        dqkv = _divided_attention_bwd(qkv_t, att_t, datt, n, 'time')
        g[b + 'timeattn.qkv.weight'], g[b + 'timeattn.qkv.bias'], dln = _lin_bwd(dqkv, ln3, W[b + 'timeattn.qkv'], True)
        _, g[b + 'norm3.weight'], g[b + 'norm3.bias'] = ops.layernorm_bwd(dln, x_in, P[b + 'norm3.weight'], EPS_V, dx=dx, accumulate=True)
    # embeddings (video_model_builder.py:221-254): token 0 = cls + pos[0]; token 1 + f*196 + s = conv3d(patch) + pos[1 + s] + temp[f]
    gsum = ops.colsum(dx.view(n, V_TOK * D)).view(V_TOK, D)
    g['cls_token'] = gsum[0].clone().view(1, 1, D)
    dpos = torch.empty((1 + V_SPACE, D), device=dx.device, dtype=torch.float32)
    dpos[0].copy_(gsum[0])
    dpos[1:].copy_(ops.colsum(gsum[1:].view(V_FRAMES, V_SPACE * D)).view(V_SPACE, D))
    g['pos_embed'] = dpos.view(1, 1 + V_SPACE, D)
    g['temp_embed'] = torch.stack([ops.colsum(gsum[1 + f * V_SPACE:1 + (f + 1) * V_SPACE]) for f in range(V_FRAMES)]).view(1, V_FRAMES, D)
    dpatch_b = ops.gather_rows_bf16(dx, n * 1568, 1568, V_TOK, 1)
    dw, g['patch_embed_3d.proj.bias'], _ = _lin_bwd(dpatch_b, sv['a'], None, False)
    g['patch_embed_3d.proj.weight'] = dw.view(D, 3, 2, 16, 16)
    return g


# ------------------------------------------------------------------------------------------------------------------
# autograd wrappers
# ------------------------------------------------------------------------------------------------------------------
class _TowerFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, names: List[str], fwd, bwd, x: torch.Tensor, *params: torch.Tensor):
        feats, saved = fwd(module, x)
        ctx.module, ctx.names, ctx.bwd, ctx.saved = module, names, bwd, saved
        return feats

    @staticmethod
    def backward(ctx, d_feats: torch.Tensor):
        saved = ctx.saved
        if saved is None:
            raise RuntimeError('backward through the encoder a second time: activations were freed after the first pass')
        ctx.saved = None
        g = ctx.bwd(ctx.module, saved, d_feats.float().contiguous())
        shapes = {n: tuple(p.shape) for n, p in ctx.module.named_parameters()}
        grads = [g[n].reshape(shapes[n]) if (ctx.needs_input_grad[5 + k] and n in g) else None for k, n in enumerate(ctx.names)]
        return (None, None, None, None, None, *grads)


def _apply(module, fwd, bwd, x: torch.Tensor) -> torch.Tensor:
    params = dict(module.named_parameters())
    return _TowerFn.apply(module, list(params), fwd, bwd, x, *params.values())


def ast_features(m, spec: torch.Tensor) -> torch.Tensor:
    """Differentiable AST.encode: spec (B, S, 128, 66) -> (B, S, 6, 768) [or (B, S, 768) with AveragePooling]."""
    ops.require_cuda(spec, 'spec')
    if spec.dim() != 4 or tuple(spec.shape[2:]) != (128, 66):
        raise ValueError(f'expected spectrogram of shape (B, S, 128, 66), got {tuple(spec.shape)}')
    B, S = spec.shape[:2]
    feats = _apply(m, _ast_forward, _ast_backward, spec.float().contiguous().view(B * S, 128, 66)).view(B, S, A_T, D)
    return _mean_tokens(feats) if m.time_pool else feats


def motionformer_features(m, vis: torch.Tensor, seed: Optional[int] = None) -> torch.Tensor:
    """Differentiable MotionFormer.encode: vis (B, S, 16, 3, 224, 224) -> (B, S, 8, 768) [or (B, S, 768)].  DropPath is active iff the
    module is in train mode."""
    ops.require_cuda(vis, 'vis')
    if vis.dim() != 6 or tuple(vis.shape[2:]) != (16, 3, 224, 224):
        raise ValueError(f'expected video of shape (B, S, 16, 3, 224, 224), got {tuple(vis.shape)}')
    B, S = vis.shape[:2]
    if seed is None:
        seed = train.draw_seed()
    stochastic = bool(m.training)
    fwd = lambda mod, x: _motionformer_forward(mod, x, seed, stochastic)
    feats = _apply(m, fwd, _motionformer_backward, vis.contiguous().view(B * S, 16, 3, 224, 224)).view(B, S, V_FRAMES, D)
    return _mean_tokens(feats) if m.time_pool else feats


def _mean_tokens(feats: torch.Tensor) -> torch.Tensor:
    from .model import _MeanTokens            # model.py imports this module: resolve at call time
    return _MeanTokens.apply(feats)
