"""Thin tensor-level wrappers over the C-ABI.  PyTorch is used for device memory and streams only.

Every function launches asynchronously on the current CUDA stream of the tensors' device and fails loudly
(`SfbError`) if the library is missing or an argument is rejected; there is no fallback implementation.
"""
import ctypes
import os
from typing import Optional

import torch

from . import _lib
from ._lib import AttnDesc, SFB_GEMM_EMIT_LN, SFB_GEMM_GELU, SFB_GEMM_LN_FOLD, SFB_GEMM_OUT_F32, SFB_GEMM_RESIDUAL, check

D = 768

# bring-up switches (tests only): SFB_GEMM_IMPL=1 / SFB_ATTN_IMPL=1 route through the plain CUDA-core kernels
GEMM_IMPL = int(os.environ.get('SFB_GEMM_IMPL', '0'))
ATTN_IMPL = int(os.environ.get('SFB_ATTN_IMPL', '0'))

_launches = 0


def launch_count() -> int:
    """Number of kernel launches issued through this module since import (bench.py reports the delta)."""
    return _launches


def _stream(t: torch.Tensor):
    """Current stream of the tensor's device.  Kernels launch on the CURRENT device, so the tensor must live there: a model on cuda:1
    without `torch.cuda.set_device(1)` / `with torch.cuda.device(1)` is refused instead of launching with a foreign stream handle."""
    if t.device.index is not None and t.device.index != torch.cuda.current_device():
        raise _lib.SfbError(f'tensor lives on {t.device} but the current CUDA device is cuda:{torch.cuda.current_device()}: '
                            'select the device first (torch.cuda.set_device / torch.cuda.device)')
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _p(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _count(n: int = 1):
    global _launches
    _launches += n


def require_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise _lib.SfbError(f'{name} must be a CUDA tensor: synchformer_b200 has no CPU path')


def device_check():
    check(_lib.load().sfb_device_check(), 'sfb_device_check')


def gemm(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], out: Optional[torch.Tensor] = None, *, gelu: bool = False,
         residual: Optional[torch.Tensor] = None, out_f32: bool = False, impl: Optional[int] = None, ln_fold=None, emit_ln=None) -> torch.Tensor:
    """out = epi(a @ w.T + bias).  a (M, K) bf16 (row stride may exceed K), w (N, K) bf16 contiguous, bias (N,) fp32,
    residual (M, N) or (1, N) fp32 (broadcast).

    LayerNorm fusion (see sfb_gemm_bf16_ln in the header):
      ln_fold = (stats (M, parts, 2) fp32, colsum (N,) fp32, eps): `a` holds UN-normalised rows, `w` / `bias` are the folded weights; the
                epilogue applies rstd (acc - mean colsum) + bias.  bf16 output.
      emit_ln = (xb (M, N) bf16, stats (M, N // 64, 2) fp32): a residual GEMM also writes the bf16 copy of its fp32 output and the per-row
                partial (sum, sum of squares) - the ln_fold inputs of the next Linear."""
    require_cuda(a, 'a')
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and a.dim() == 2 and w.dim() == 2
    assert a.stride(1) == 1 and w.is_contiguous()
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K
    if out is None:
        out = torch.empty((M, N), device=a.device, dtype=torch.float32 if out_f32 else torch.bfloat16)
    assert out.shape == (M, N) and out.stride(1) == 1 and out.dtype == (torch.float32 if out_f32 else torch.bfloat16)
    flags = (SFB_GEMM_GELU if gelu else 0) | (SFB_GEMM_OUT_F32 if out_f32 else 0)
    ldr = 0
    if residual is not None:
        assert residual.dtype == torch.float32 and residual.stride(-1) == 1 and residual.shape[-1] == N
        flags |= SFB_GEMM_RESIDUAL
        ldr = 0 if residual.numel() == N else residual.stride(0)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == N and bias.is_contiguous()
    impl = GEMM_IMPL if impl is None else impl
    if ln_fold is None and emit_ln is None:
        check(_lib.load().sfb_gemm_bf16(_p(a), a.stride(0), _p(w), _p(bias), _p(residual), ldr, _p(out), out.stride(0), M, N, K, flags,
                                        impl, _stream(a)), 'sfb_gemm_bf16')
        _count()
        return out
    stats_in = colsum = stats_out = xb = None
    parts, eps, ld_emit = 0, 0.0, 0
    if ln_fold is not None:
        stats_in, colsum, eps = ln_fold
        assert stats_in.dtype == torch.float32 and stats_in.is_contiguous() and stats_in.dim() == 3 and stats_in.shape[0] >= M and stats_in.shape[2] == 2
        assert colsum.dtype == torch.float32 and colsum.is_contiguous() and colsum.numel() == N and bias is not None and not out_f32
        parts = stats_in.shape[1]
        flags |= SFB_GEMM_LN_FOLD
    if emit_ln is not None:
        xb, stats_out = emit_ln
        assert out_f32 and N % 64 == 0 and xb.dtype == torch.bfloat16 and xb.shape == (M, N) and xb.stride(1) == 1
        assert stats_out.dtype == torch.float32 and stats_out.is_contiguous() and tuple(stats_out.shape) == (M, N // 64, 2)
        ld_emit = xb.stride(0)
        flags |= SFB_GEMM_EMIT_LN
        if impl == 1:
            raise _lib.SfbError('emit_ln is not available on the CUDA-core cross-check kernel')
    check(_lib.load().sfb_gemm_bf16_ln(_p(a), a.stride(0), _p(w), _p(bias), _p(residual), ldr, _p(out), out.stride(0), M, N, K, flags, impl,
                                       _p(stats_in), parts, _p(colsum), float(eps), _p(stats_out), _p(xb), ld_emit, _stream(a)), 'sfb_gemm_bf16_ln')
    _count()
    return out


def rowstats_cast(x: torch.Tensor, xb: Optional[torch.Tensor] = None, stats: Optional[torch.Tensor] = None):
    """x (R, 768) fp32 -> (xb (R, 768) bf16, stats (R, 1, 2) fp32 = per-row (sum, sum of squares)): the `ln_fold` inputs of gemm() for rows
    that did not come out of an `emit_ln` GEMM."""
    require_cuda(x, 'x')
    assert x.dtype == torch.float32 and x.dim() == 2 and x.shape[1] == D and x.stride(1) == 1
    R = x.shape[0]
    if xb is None:
        xb = torch.empty((R, D), device=x.device, dtype=torch.bfloat16)
    if stats is None:
        stats = torch.empty((R, 1, 2), device=x.device, dtype=torch.float32)
    assert xb.dtype == torch.bfloat16 and xb.is_contiguous() and xb.shape == (R, D) and stats.is_contiguous() and tuple(stats.shape) == (R, 1, 2)
    check(_lib.load().sfb_rowstats_cast(_p(x), x.stride(0), _p(xb), _p(stats), R, _stream(x)), 'sfb_rowstats_cast')
    _count()
    return xb, stats


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float, out: Optional[torch.Tensor] = None, *,
              rows: Optional[int] = None, group: Optional[int] = None, group_stride: Optional[int] = None, offset: int = 0,
              gamma2: Optional[torch.Tensor] = None, beta2: Optional[torch.Tensor] = None, eps2: float = 0.0,
              out_f32: bool = False) -> torch.Tensor:
    """Row LayerNorm over 768 of fp32 x (R, 768).  Output row r reads input row (r // group) * group_stride + offset + r % group."""
    require_cuda(x, 'x')
    assert x.dtype == torch.float32 and x.dim() == 2 and x.shape[1] == D and x.stride(1) == 1
    if rows is None:
        rows = x.shape[0]
    if group is None:
        group, group_stride = rows, rows
    assert ((rows - 1) // group) * group_stride + offset + (rows - 1) % group < x.shape[0], 'row gather out of range'
    if out is None:
        out = torch.empty((rows, D), device=x.device, dtype=torch.float32 if out_f32 else torch.bfloat16)
    assert out.shape[0] >= rows and out.shape[1] == D and out.stride(1) == 1
    assert out.dtype == (torch.float32 if out_f32 else torch.bfloat16)
    check(_lib.load().sfb_layernorm(_p(x), x.stride(0), _p(out), out.stride(0), int(out_f32), _p(gamma), _p(beta), eps, _p(gamma2), _p(beta2),
                                    eps2, rows, group, group_stride, offset, _stream(x)), 'sfb_layernorm')
    _count()
    return out


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, out: torch.Tensor, *, q_strides, kv_strides, o_strides, n_outer: int,
              n_inner: int, n_heads: int, head_dim: int, Lq: int, Lk: int, scale: float, k_prefix: Optional[torch.Tensor] = None,
              v_prefix: Optional[torch.Tensor] = None, prefix_outer: int = 0, impl: Optional[int] = None,
              q_extra: Optional[torch.Tensor] = None, q_extra_outer: int = 0, extra_out: Optional[torch.Tensor] = None,
              extra_out_outer: int = 0) -> bool:
    """softmax(scale q k^T) v on strided bf16 views; q/k/v/out are tensors whose data_ptr() is the address of problem
    (0, 0), head 0, row 0; *_strides = (outer, inner, row) in elements.  See sfb_attn_desc in the header.

    q_extra / extra_out: ask for the fused extra query (one more query row per (outer, head) that attends to the keys of ALL inner
    problems of its outer index, e.g. the Motionformer CLS query); its merged result goes to extra_out + o*extra_out_outer + h*head_dim.
    Returns True if the extra query was fused; False if this descriptor does not support it (then only the regular rows were computed
    and the caller runs the extra query as its own attention call)."""
    require_cuda(q, 'q')
    for t in (q, k, v, out):
        assert t.dtype == torch.bfloat16
    d = AttnDesc()
    d.q, d.k, d.v, d.out = q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr()
    d.k_prefix = None if k_prefix is None else k_prefix.data_ptr()
    d.v_prefix = None if v_prefix is None else v_prefix.data_ptr()
    d.q_outer, d.q_inner, d.q_row = q_strides
    d.kv_outer, d.kv_inner, d.kv_row = kv_strides
    d.o_outer, d.o_inner, d.o_row = o_strides
    d.prefix_outer = prefix_outer
    d.n_outer, d.n_inner, d.n_heads, d.head_dim, d.Lq, d.Lk = n_outer, n_inner, n_heads, head_dim, Lq, Lk
    d.scale = scale
    d.impl = ATTN_IMPL if impl is None else impl
    d.q_extra, d.q_extra_outer, d.extra_partial = None, 0, None
    lib = _lib.load()
    fused, partial = False, None
    if q_extra is not None:
        assert extra_out is not None and q_extra.dtype == torch.bfloat16 and extra_out.dtype == torch.bfloat16
        d.q_extra, d.q_extra_outer = q_extra.data_ptr(), q_extra_outer
        if lib.sfb_attention_extra_supported(ctypes.byref(d)) == 1:
            partial = torch.empty((n_outer * n_heads * n_inner, head_dim + 2), device=q.device, dtype=torch.float32)
            d.extra_partial = partial.data_ptr()
            fused = True
        else:
            d.q_extra, d.q_extra_outer = None, 0
    check(lib.sfb_attention(ctypes.byref(d), _stream(q)), 'sfb_attention')
    _count()
    if fused:
        check(lib.sfb_attention_merge_partials(_p(partial), _p(extra_out), extra_out_outer, n_outer, n_inner, n_heads, head_dim, _stream(q)),
              'sfb_attention_merge_partials')
        _count()
    return fused


_VIDEO_DTYPES = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2, torch.uint8: 3}


def im2col_video(vis: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """vis (n_seg, 16, 3, 224, 224) fp32 / fp16 / bf16 / uint8 contiguous -> (n_seg * 1568, 1536) bf16."""
    require_cuda(vis, 'vis')
    assert vis.is_contiguous() and tuple(vis.shape[1:]) == (16, 3, 224, 224), f'bad video shape {tuple(vis.shape)}'
    if vis.dtype not in _VIDEO_DTYPES:
        raise _lib.SfbError(f'unsupported video dtype {vis.dtype}')
    n = vis.shape[0]
    if out is None:
        out = torch.empty((n * 1568, 1536), device=vis.device, dtype=torch.bfloat16)
    check(_lib.load().sfb_im2col_video(_p(vis), _VIDEO_DTYPES[vis.dtype], _p(out), n, _stream(vis)), 'sfb_im2col_video')
    _count()
    return out


def im2col_video_clip(clip: torch.Tensor, n_segments: int, v_start: int, v_stride: int) -> torch.Tensor:
    """clip (n_clips, n_frames, 3, 224, 224) -> (n_clips * n_segments * 1568, 1536) bf16; segment s = frames [v_start + s*v_stride, +16)."""
    require_cuda(clip, 'clip')
    assert clip.is_contiguous() and clip.dim() == 5 and tuple(clip.shape[2:]) == (3, 224, 224), f'bad clip shape {tuple(clip.shape)}'
    if clip.dtype not in _VIDEO_DTYPES:
        raise _lib.SfbError(f'unsupported video dtype {clip.dtype}')
    n_clips, n_frames = clip.shape[:2]
    out = torch.empty((n_clips * n_segments * 1568, 1536), device=clip.device, dtype=torch.bfloat16)
    check(_lib.load().sfb_im2col_video_clip(_p(clip), _VIDEO_DTYPES[clip.dtype], _p(out), n_clips, n_frames, n_segments, v_start, v_stride,
                                            _stream(clip)), 'sfb_im2col_video_clip')
    _count()
    return out


def video_tokens(patch: torch.Tensor, cls_token: torch.Tensor, pos_embed: torch.Tensor, temp_embed: torch.Tensor, n_seg: int,
                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
    if out is None:
        out = torch.empty((n_seg * 1569, D), device=patch.device, dtype=torch.float32)
    check(_lib.load().sfb_video_tokens(_p(patch), _p(cls_token), _p(pos_embed), _p(temp_embed), _p(out), n_seg, _stream(patch)),
          'sfb_video_tokens')
    _count()
    return out


def im2col_ast(spec: torch.Tensor) -> torch.Tensor:
    """spec (n_seg, 128, 66) fp32 contiguous [freq, time] -> (n_seg * 72, 256) bf16."""
    require_cuda(spec, 'spec')
    assert spec.dtype == torch.float32 and spec.is_contiguous() and tuple(spec.shape[1:]) == (128, 66), f'bad mel shape {tuple(spec.shape)}'
    n = spec.shape[0]
    out = torch.empty((n * 72, 256), device=spec.device, dtype=torch.bfloat16)
    check(_lib.load().sfb_im2col_ast(_p(spec), _p(out), n, _stream(spec)), 'sfb_im2col_ast')
    _count()
    return out


def ast_tokens(patch: torch.Tensor, cls_token: torch.Tensor, dist_token: torch.Tensor, pos_embed: torch.Tensor, n_seg: int) -> torch.Tensor:
    out = torch.empty((n_seg * 74, D), device=patch.device, dtype=torch.float32)
    check(_lib.load().sfb_ast_tokens(_p(patch), _p(cls_token), _p(dist_token), _p(pos_embed), _p(out), n_seg, _stream(patch)), 'sfb_ast_tokens')
    _count()
    return out


def sync_tokens(v: torch.Tensor, a: torch.Tensor, vis_ln_w, vis_ln_b, aud_ln_w, aud_ln_b, eps: float, off_tok, mod_tok, pos_emb, B: int,
                S: int) -> torch.Tensor:
    out = torch.empty((B * (2 + 14 * S), D), device=v.device, dtype=torch.float32)
    check(_lib.load().sfb_sync_tokens(_p(v), _p(a), _p(vis_ln_w), _p(vis_ln_b), _p(aud_ln_w), _p(aud_ln_b), eps, _p(off_tok), _p(mod_tok),
                                      _p(pos_emb), _p(out), B, S, _stream(v)), 'sfb_sync_tokens')
    _count()
    return out


def sync_head(x: torch.Tensor, T: int, ln_w, ln_b, eps: float, W: torch.Tensor, b: torch.Tensor, B: int) -> torch.Tensor:
    n_cls = W.shape[0]
    out = torch.empty((B, n_cls), device=x.device, dtype=torch.float32)
    check(_lib.load().sfb_sync_head(_p(x), T, _p(ln_w), _p(ln_b), eps, _p(W), _p(b), _p(out), B, n_cls, _stream(x)), 'sfb_sync_head')
    _count()
    return out


def cast_bf16(x: torch.Tensor) -> torch.Tensor:
    require_cuda(x, 'x')
    assert x.dtype == torch.float32 and x.is_contiguous() and x.numel() % 4 == 0
    out = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    check(_lib.load().sfb_cast_f32_bf16(_p(x), _p(out), x.numel(), _stream(x)), 'sfb_cast_f32_bf16')
    _count()
    return out


def mel_frontend(wave: torch.Tensor) -> torch.Tensor:
    """wave (..., 10240) fp32 -> normalised log-mel (..., 128, 66) fp32 (dataset/transforms.py:815-871)."""
    require_cuda(wave, 'wave')
    assert wave.dtype == torch.float32 and wave.is_contiguous() and wave.shape[-1] == 10240
    n = wave.numel() // 10240
    out = torch.empty((*wave.shape[:-1], 128, 66), device=wave.device, dtype=torch.float32)
    check(_lib.load().sfb_mel_frontend(_p(wave), _p(out), n, _stream(wave)), 'sfb_mel_frontend')
    _count()
    return out


def mel_frontend_clip(wave: torch.Tensor, n_segments: int, a_start: int, a_stride: int) -> torch.Tensor:
    """wave (n_clips, n_samples) fp32 un-duplicated waveforms -> (n_clips, n_segments, 128, 66); segment s = samples [a_start + s*a_stride, +10240)."""
    require_cuda(wave, 'wave')
    assert wave.dtype == torch.float32 and wave.is_contiguous() and wave.dim() == 2
    n_clips = wave.shape[0]
    out = torch.empty((n_clips, n_segments, 128, 66), device=wave.device, dtype=torch.float32)
    check(_lib.load().sfb_mel_frontend_clip(_p(wave), wave.shape[1], _p(out), n_clips, n_segments, a_start, a_stride, _stream(wave)),
          'sfb_mel_frontend_clip')
    _count()
    return out


# ------------------------------------------------------------------------------------------------------------------
# N3: training kernels of the synchronisation module (see include/synchformer_b200.h, "N3")
# ------------------------------------------------------------------------------------------------------------------
def dropout(x: torch.Tensor, p: float, seed: int, site: int, *, residual: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
            out_bf16: bool = False) -> torch.Tensor:
    """out = residual + x * mask / (1 - p) with the counter-based mask keep(seed, site, linear index); x / residual fp32."""
    require_cuda(x, 'x')
    assert x.dtype == torch.float32 and x.is_contiguous() and x.numel() % 4 == 0
    if residual is not None:
        assert residual.dtype == torch.float32 and residual.is_contiguous() and residual.shape == x.shape
    if out is None:
        out = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16 if out_bf16 else torch.float32)
    assert out.is_contiguous() and out.shape == x.shape and out.dtype == (torch.bfloat16 if out_bf16 else torch.float32)
    check(_lib.load().sfb_dropout(_p(x), _p(residual), _p(out), int(out_bf16), x.numel(), float(p), int(seed), int(site), _stream(x)), 'sfb_dropout')
    _count()
    return out


def gelu_fwd(x: torch.Tensor) -> torch.Tensor:
    require_cuda(x, 'x')
    assert x.dtype == torch.bfloat16 and x.is_contiguous() and x.numel() % 8 == 0
    y = torch.empty_like(x)
    check(_lib.load().sfb_gelu_fwd(_p(x), _p(y), x.numel(), _stream(x)), 'sfb_gelu_fwd')
    _count()
    return y


def gelu_bwd(dy: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    require_cuda(x, 'x')
    assert x.dtype == torch.bfloat16 and dy.dtype == torch.bfloat16 and x.is_contiguous() and dy.is_contiguous() and x.shape == dy.shape
    assert x.numel() % 8 == 0
    dx = torch.empty_like(x)
    check(_lib.load().sfb_gelu_bwd(_p(dy), _p(x), _p(dx), x.numel(), _stream(x)), 'sfb_gelu_bwd')
    _count()
    return dx


def transpose_bf16(x: torch.Tensor, pad_to: int = 8) -> torch.Tensor:
    """(R, C) bf16 (row stride may exceed C) -> (C, ceil(R / pad_to) * pad_to) bf16 whose padding columns are zero (pad_to % 8 == 0:
    the result is used as a GEMM operand with K = its width)."""
    require_cuda(x, 'x')
    assert x.dtype == torch.bfloat16 and x.dim() == 2 and x.stride(1) == 1 and pad_to % 8 == 0
    R, C = x.shape
    ld = (R + pad_to - 1) // pad_to * pad_to
    out = torch.empty((C, ld), device=x.device, dtype=torch.bfloat16)
    check(_lib.load().sfb_transpose_bf16(_p(x), x.stride(0), R, C, _p(out), ld, _stream(x)), 'sfb_transpose_bf16')
    _count()
    return out


def colsum(x: torch.Tensor) -> torch.Tensor:
    """(M, N) bf16 / fp32 -> (N,) fp32 column sums (deterministic)."""
    require_cuda(x, 'x')
    assert x.dtype in (torch.bfloat16, torch.float32) and x.dim() == 2 and x.stride(1) == 1
    M, N = x.shape
    out = torch.empty((N,), device=x.device, dtype=torch.float32)
    parts = min(64, (M + 63) // 64)
    ws = torch.empty((parts * N,), device=x.device, dtype=torch.float32) if parts > 1 else None
    check(_lib.load().sfb_colsum(_p(x), int(x.dtype == torch.bfloat16), x.stride(0), M, N, _p(out), _p(ws), 0 if ws is None else ws.numel(),
                                 _stream(x)), 'sfb_colsum')
    _count(2 if parts > 1 else 1)
    return out


def layernorm_bwd(dy: torch.Tensor, x: torch.Tensor, gamma: torch.Tensor, eps: float, *, dx: Optional[torch.Tensor] = None,
                  accumulate: bool = False, rows: Optional[int] = None, group: Optional[int] = None, group_stride: Optional[int] = None,
                  offset: int = 0):
    """LayerNorm backward over 768 (fp32).  x (rows, 768); dy row of output row r = (r // group) * group_stride + offset + r % group.
    Returns (dx (rows, 768), dgamma (768,), dbeta (768,)); accumulate=True adds into the given dx."""
    require_cuda(x, 'x')
    assert dy.dtype == torch.float32 and x.dtype == torch.float32 and dy.dim() == 2 and x.dim() == 2 and dy.stride(1) == 1 and x.stride(1) == 1
    assert x.shape[1] == D and dy.shape[1] == D
    if rows is None:
        rows = x.shape[0]
    if group is None:
        group, group_stride = rows, rows
    assert ((rows - 1) // group) * group_stride + offset + (rows - 1) % group < dy.shape[0], 'dy row gather out of range'
    assert x.shape[0] >= rows
    if dx is None:
        assert not accumulate
        dx = torch.empty((rows, D), device=x.device, dtype=torch.float32)
    assert dx.dtype == torch.float32 and dx.shape[0] >= rows and dx.shape[1] == D and dx.stride(1) == 1
    lib = _lib.load()
    dgb = torch.empty((2, D), device=x.device, dtype=torch.float32)
    n_ws = lib.sfb_layernorm_bwd_workspace_floats(rows)
    ws = torch.empty((n_ws,), device=x.device, dtype=torch.float32)
    check(lib.sfb_layernorm_bwd(_p(dy), dy.stride(0), group, group_stride, offset, _p(x), x.stride(0), _p(gamma), eps, _p(dx), dx.stride(0),
                                int(accumulate), _p(dgb[0]), _p(dgb[1]), _p(ws), n_ws, rows, _stream(x)), 'sfb_layernorm_bwd')
    _count(2)
    return dx, dgb[0], dgb[1]


def attention_train_fwd(qkv: torch.Tensor, B: int, T: int, n_heads: int, head_dim: int, scale: float, p: float, seed: int, site: int):
    """qkv (B*T, 3*n_heads*head_dim) bf16 -> (out (B*T, n_heads*head_dim) bf16, lse (B, n_heads, T) fp32)."""
    require_cuda(qkv, 'qkv')
    Dm = n_heads * head_dim
    assert qkv.dtype == torch.bfloat16 and qkv.is_contiguous() and qkv.shape == (B * T, 3 * Dm)
    out = torch.empty((B * T, Dm), device=qkv.device, dtype=torch.bfloat16)
    lse = torch.empty((B, n_heads, T), device=qkv.device, dtype=torch.float32)
    check(_lib.load().sfb_attention_train_fwd(_p(qkv), _p(out), _p(lse), B, T, n_heads, head_dim, scale, float(p), int(seed), int(site),
                                              _stream(qkv)), 'sfb_attention_train_fwd')
    _count()
    return out, lse


def attention_train_bwd(qkv: torch.Tensor, out: torch.Tensor, d_out: torch.Tensor, lse: torch.Tensor, B: int, T: int, n_heads: int,
                        head_dim: int, scale: float, p: float, seed: int, site: int) -> torch.Tensor:
    """-> dqkv (B*T, 3*n_heads*head_dim) bf16."""
    require_cuda(qkv, 'qkv')
    Dm = n_heads * head_dim
    assert qkv.dtype == torch.bfloat16 and qkv.is_contiguous() and qkv.shape == (B * T, 3 * Dm)
    for t in (out, d_out):
        assert t.dtype == torch.bfloat16 and t.is_contiguous() and t.shape == (B * T, Dm)
    assert lse.dtype == torch.float32 and lse.is_contiguous() and lse.numel() == B * n_heads * T
    dqkv = torch.empty_like(qkv)
    delta = torch.empty_like(lse)
    check(_lib.load().sfb_attention_train_bwd(_p(qkv), _p(out), _p(d_out), _p(lse), _p(delta), _p(dqkv), B, T, n_heads, head_dim, scale, float(p),
                                              int(seed), int(site), _stream(qkv)), 'sfb_attention_train_bwd')
    _count(2)
    return dqkv


def sync_head_bwd(x: torch.Tensor, T: int, ln_w: torch.Tensor, ln_b: torch.Tensor, eps: float, W: torch.Tensor, dlogits: torch.Tensor, B: int):
    """backward of sync_head: -> (dx (B*T, 768) fp32 [zero except token 0 of each clip], dln_w, dln_b, dW (n_cls, 768), dbias (n_cls))."""
    require_cuda(x, 'x')
    n_cls = W.shape[0]
    assert x.dtype == torch.float32 and x.is_contiguous() and x.shape == (B * T, D)
    assert dlogits.dtype == torch.float32 and dlogits.is_contiguous() and dlogits.shape == (B, n_cls)
    assert W.dtype == torch.float32 and W.is_contiguous()
    dev = x.device
    dx = torch.empty((B * T, D), device=dev, dtype=torch.float32)
    dln = torch.empty((2, D), device=dev, dtype=torch.float32)
    dW = torch.empty((n_cls, D), device=dev, dtype=torch.float32)
    db = torch.empty((n_cls,), device=dev, dtype=torch.float32)
    scratch = torch.empty((B * 3 * D,), device=dev, dtype=torch.float32)
    check(_lib.load().sfb_sync_head_bwd(_p(x), T, _p(ln_w), _p(ln_b), eps, _p(W), _p(dlogits), B, n_cls, _p(dx), _p(dln[0]), _p(dln[1]), _p(dW),
                                        _p(db), _p(scratch), _stream(x)), 'sfb_sync_head_bwd')
    _count(2)
    return dx, dln[0], dln[1], dW, db


# ------------------------------------------------------------------------------------------------------------------
# N1: backward of the encoders (see include/synchformer_b200.h, "N1")
# ------------------------------------------------------------------------------------------------------------------
def empty_bf16(shape, device) -> torch.Tensor:
    """activation buffer (one place to allocate them, so tests can swap the storage type of the stand-ins)"""
    return torch.empty(shape, device=device, dtype=torch.bfloat16)


def attention_bwd(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, out: torch.Tensor, d_out: torch.Tensor, dq: torch.Tensor, dk: torch.Tensor,
                  dv: torch.Tensor, *, q_strides, kv_strides, o_strides, n_outer: int, n_inner: int, n_heads: int, head_dim: int, Lq: int, Lk: int,
                  scale: float, k_prefix: Optional[torch.Tensor] = None, v_prefix: Optional[torch.Tensor] = None, prefix_outer: int = 0,
                  impl: int = 0):
    """Backward of `attention(...)` with the same view arguments; d_out is addressed like out, dq / dk / dv like q / k / v.
    impl: 0 = auto (mma.sync kernel for head_dim 64 and 64 <= Lq, Lk + prefix <= 256; CUDA-core kernels otherwise), 1 = CUDA-core kernels.
    Returns the per-problem prefix gradients (n_inner, n_outer, n_heads, 2, head_dim) fp32 (None without a prefix): sum them over the
    problems that share the prefix row (colsum)."""
    require_cuda(q, 'q')
    for t in (q, k, v, out, d_out, dq, dk, dv):
        assert t.dtype == torch.bfloat16
    d = AttnDesc()
    d.q, d.k, d.v, d.out = q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr()
    d.k_prefix = None if k_prefix is None else k_prefix.data_ptr()
    d.v_prefix = None if v_prefix is None else v_prefix.data_ptr()
    d.q_outer, d.q_inner, d.q_row = q_strides
    d.kv_outer, d.kv_inner, d.kv_row = kv_strides
    d.o_outer, d.o_inner, d.o_row = o_strides
    d.prefix_outer = prefix_outer
    d.n_outer, d.n_inner, d.n_heads, d.head_dim, d.Lq, d.Lk = n_outer, n_inner, n_heads, head_dim, Lq, Lk
    d.scale, d.impl = scale, impl
    d.q_extra, d.q_extra_outer, d.extra_partial = None, 0, None
    lib = _lib.load()
    stats = torch.empty((lib.sfb_attention_bwd_stats_floats(ctypes.byref(d)),), device=q.device, dtype=torch.float32)
    dprefix = None if k_prefix is None else torch.empty((n_inner, n_outer, n_heads, 2, head_dim), device=q.device, dtype=torch.float32)
    check(lib.sfb_attention_bwd(ctypes.byref(d), _p(d_out), _p(dq), _p(dk), _p(dv), _p(dprefix), _p(stats), _stream(q)), 'sfb_attention_bwd')
    _count(2)
    return dprefix


def attention_bwd_global_query(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, out: torch.Tensor, d_out: torch.Tensor, dq: torch.Tensor,
                               dk: torch.Tensor, dv: torch.Tensor, *, q_outer: int, kv_outer: int, kv_row: int, o_outer: int, n_outer: int, n_heads: int,
                               head_dim: int, Lk: int, scale: float, prefix_grad: Optional[torch.Tensor] = None):
    """Backward of the one-query-per-(outer, head) attention over all Lk rows: writes dq, ADDS to dk / dv (and prefix_grad to row 0)."""
    require_cuda(q, 'q')
    for t in (q, k, v, out, d_out, dq, dk, dv):
        assert t.dtype == torch.bfloat16
    if prefix_grad is not None:
        assert prefix_grad.dtype == torch.float32 and prefix_grad.is_contiguous() and prefix_grad.numel() == n_outer * n_heads * 2 * head_dim
    coef = torch.empty((n_outer * n_heads * Lk * 2,), device=q.device, dtype=torch.float32)
    check(_lib.load().sfb_attention_bwd_global_query(_p(q), q_outer, _p(k), _p(v), kv_outer, kv_row, _p(out), _p(d_out), o_outer, _p(dq), _p(dk), _p(dv),
                                                     _p(prefix_grad), _p(coef), n_outer, n_heads, head_dim, Lk, scale, _stream(q)),
          'sfb_attention_bwd_global_query')
    _count(2)


def droppath(x: torch.Tensor, rows_per_sample: int, p: float, seed: int, site: int, *, residual: Optional[torch.Tensor] = None,
             out: Optional[torch.Tensor] = None, out_bf16: bool = False) -> torch.Tensor:
    """out = residual + x * keep(sample) / (1 - p), one decision per `rows_per_sample` consecutive rows; x / residual (rows, 768) fp32."""
    require_cuda(x, 'x')
    assert x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 2 and x.shape[1] == D and x.shape[0] % rows_per_sample == 0
    if residual is not None:
        assert residual.dtype == torch.float32 and residual.is_contiguous() and residual.shape == x.shape
    if out is None:
        out = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16 if out_bf16 else torch.float32)
    assert out.is_contiguous() and out.shape == x.shape and out.dtype == (torch.bfloat16 if out_bf16 else torch.float32)
    check(_lib.load().sfb_droppath(_p(x), _p(residual), _p(out), int(out_bf16), x.shape[0], rows_per_sample, float(p), int(seed), int(site), _stream(x)),
          'sfb_droppath')
    _count()
    return out


def gather_rows_bf16(x: torch.Tensor, rows: int, group: Optional[int] = None, group_stride: Optional[int] = None, offset: int = 0) -> torch.Tensor:
    """(R, 768) fp32 -> (rows, 768) bf16; output row r reads input row (r // group) * group_stride + offset + r % group."""
    require_cuda(x, 'x')
    assert x.dtype == torch.float32 and x.dim() == 2 and x.shape[1] == D and x.stride(1) == 1
    if group is None:
        group, group_stride = rows, rows
    assert ((rows - 1) // group) * group_stride + offset + (rows - 1) % group < x.shape[0], 'row gather out of range'
    out = torch.empty((rows, D), device=x.device, dtype=torch.bfloat16)
    check(_lib.load().sfb_gather_rows_bf16(_p(x), x.stride(0), _p(out), rows, group, group_stride, offset, _stream(x)), 'sfb_gather_rows_bf16')
    _count()
    return out


# ------------------------------------------------------------------------------------------------------------------
# N1: tail of the stage-I contrastive step (see include/synchformer_b200.h, "contrastive")
# ------------------------------------------------------------------------------------------------------------------
def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    require_cuda(t, name)
    assert t.dtype == torch.float32 and t.is_contiguous(), f'{name} must be contiguous fp32'
    return t


def mean_tokens(x: torch.Tensor) -> torch.Tensor:
    """(n, T, D) fp32 -> (n, D): AveragePooling over the token rows (motionformer.py:405-409)."""
    _f32c(x, 'x')
    n, T, Dm = x.shape
    out = torch.empty((n, Dm), device=x.device, dtype=torch.float32)
    check(_lib.load().sfb_mean_tokens(_p(x), _p(out), n, T, Dm, _stream(x)), 'sfb_mean_tokens')
    _count()
    return out


def mean_tokens_bwd(dout: torch.Tensor, T: int) -> torch.Tensor:
    _f32c(dout, 'dout')
    n, Dm = dout.shape
    dx = torch.empty((n, T, Dm), device=dout.device, dtype=torch.float32)
    check(_lib.load().sfb_mean_tokens_bwd(_p(dout), _p(dx), n, T, Dm, _stream(dout)), 'sfb_mean_tokens_bwd')
    _count()
    return dx


def l2_normalize(x: torch.Tensor):
    """F.normalize(x, dim=-1) on (n, D) fp32 -> (xn, inv_norm (n,))."""
    _f32c(x, 'x')
    n, Dm = x.shape
    xn = torch.empty_like(x)
    inv = torch.empty((n,), device=x.device, dtype=torch.float32)
    check(_lib.load().sfb_l2_normalize(_p(x), _p(xn), _p(inv), n, Dm, _stream(x)), 'sfb_l2_normalize')
    _count()
    return xn, inv


def l2_normalize_bwd(xn: torch.Tensor, inv_norm: torch.Tensor, dxn: torch.Tensor) -> torch.Tensor:
    _f32c(xn, 'xn'), _f32c(inv_norm, 'inv_norm'), _f32c(dxn, 'dxn')
    n, Dm = xn.shape
    dx = torch.empty_like(xn)
    check(_lib.load().sfb_l2_normalize_bwd(_p(xn), _p(inv_norm), _p(dxn), _p(dx), n, Dm, _stream(xn)), 'sfb_l2_normalize_bwd')
    _count()
    return dx


def contrastive_loss(vn: torch.Tensor, an: torch.Tensor, vn_all: torch.Tensor, an_all: torch.Tensor, scale: torch.Tensor):
    """open_clip/model.py:507-527 on normalised features: local rows (n, D) as queries, gathered rows (N, D) as keys; scale: the 1-element
    fp32 device tensor holding logit_scale (read on the device).  -> (loss (1,), dscale (1,), G (2, n, N)) device tensors."""
    for t, nm in ((vn, 'vn'), (an, 'an'), (vn_all, 'vn_all'), (an_all, 'an_all'), (scale, 'scale')):
        _f32c(t, nm)
    n, Dm = vn.shape
    N = vn_all.shape[0]
    assert an.shape == (n, Dm) and vn_all.shape == (N, Dm) and an_all.shape == (N, Dm) and scale.numel() == 1
    dev = vn.device
    loss, dscale = torch.empty((1,), device=dev, dtype=torch.float32), torch.empty((1,), device=dev, dtype=torch.float32)
    G = torch.empty((2, n, N), device=dev, dtype=torch.float32)
    ws = torch.empty((4 * n,), device=dev, dtype=torch.float32)
    check(_lib.load().sfb_contrastive_loss(_p(vn), _p(an), _p(vn_all), _p(an_all), n, N, Dm, _p(scale), _p(loss), _p(dscale), _p(G), _p(ws), _stream(vn)),
          'sfb_contrastive_loss')
    _count(2)
    return loss, dscale, G


def contrastive_loss_bwd(vn: torch.Tensor, an: torch.Tensor, vn_all: Optional[torch.Tensor], an_all: Optional[torch.Tensor], G: torch.Tensor,
                         scale: torch.Tensor, upstream: torch.Tensor, dscale: torch.Tensor):
    """-> (d_vn, d_an (n, D), d_vn_all, d_an_all (N, D) or None, dscale_out (1,)), all times the device scalar `upstream`.
    vn_all / an_all None: the keys are the local rows themselves (no gathering) and their gradient is folded into d_vn / d_an."""
    n, Dm = vn.shape
    local = vn_all is None
    kv, ka = (vn, an) if local else (vn_all, an_all)
    N = kv.shape[0]
    for t, nm in ((G, 'G'), (scale, 'scale'), (upstream, 'upstream'), (dscale, 'dscale')):
        _f32c(t, nm)
    d_vn, d_an = torch.empty_like(vn), torch.empty_like(an)
    d_vn_all, d_an_all = (None, None) if local else (torch.empty_like(kv), torch.empty_like(ka))
    dscale_out = torch.empty((1,), device=vn.device, dtype=torch.float32)
    check(_lib.load().sfb_contrastive_loss_bwd(_p(vn), _p(an), _p(kv), _p(ka), _p(G), n, N, Dm, _p(scale), _p(upstream), _p(dscale), _p(d_vn), _p(d_an),
                                               _p(d_vn_all), _p(d_an_all), _p(dscale_out), _stream(vn)), 'sfb_contrastive_loss_bwd')
    _count(2)
    return d_vn, d_an, d_vn_all, d_an_all, dscale_out
