"""GPU: every kernel of the C-ABI against a plain PyTorch fp32 restatement of the same op on the same (bf16-rounded) inputs."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _bf(t):
    return t.to(torch.bfloat16)


def rel_l2(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


# ------------------------------------------------------------------------------------------------------ GEMM
GEMM_SHAPES = [
    # M, N, K            (tails in M, N and K; 1-row; every (N, K) the model uses)
    (1, 2304, 768), (77, 768, 768), (128, 256, 64), (300, 768, 768), (1000, 2304, 768), (1569 * 2, 768, 3072),
    (513, 3072, 768), (3137, 768, 1536), (144, 768, 256), (260, 1536, 768), (130, 40, 72), (4096, 768, 768),
]


@pytest.mark.parametrize('impl', [pytest.param(0, id='tc'), pytest.param(2, id='tc1cta'), pytest.param(3, id='tcpair'), pytest.param(1, id='simple')])
@pytest.mark.parametrize('M,N,K', GEMM_SHAPES)
def test_gemm_bias(cuda_device, impl, M, N, K):
    from synchformer_b200 import ops
    g = torch.Generator(device='cuda').manual_seed(M * 7 + N * 3 + K)
    a = _bf(torch.randn(M, K, device='cuda', generator=g))
    w = _bf(torch.randn(N, K, device='cuda', generator=g) * 0.05)
    b = torch.randn(N, device='cuda', generator=g)
    out = ops.gemm(a, w, b, impl=impl)
    ref = a.float() @ w.float().T + b
    torch.cuda.synchronize()
    assert out.dtype == torch.bfloat16
    assert rel_l2(out.float(), ref) < 4e-3          # bf16 output rounding: 2^-9 relative per element
    assert (out.float() - ref).abs().max() <= 1e-2 * ref.abs().max() + 1e-3


@pytest.mark.parametrize('impl', [pytest.param(0, id='tc'), pytest.param(2, id='tc1cta'), pytest.param(1, id='simple')])
def test_gemm_epilogues(cuda_device, impl):
    from synchformer_b200 import ops
    g = torch.Generator(device='cuda').manual_seed(3)
    M, N, K = 1000, 768, 768
    a = _bf(torch.randn(M, K, device='cuda', generator=g))
    w = _bf(torch.randn(N, K, device='cuda', generator=g) * 0.05)
    b = torch.randn(N, device='cuda', generator=g)
    res = torch.randn(M, N, device='cuda', generator=g)
    lin = a.float() @ w.float().T + b
    # fp32 output, no rounding of the result: tight tolerance pins accumulation + bias
    out = ops.gemm(a, w, b, out_f32=True, impl=impl)
    assert rel_l2(out, lin) < 2e-6
    out = ops.gemm(a, w, None, out_f32=True, impl=impl)
    assert rel_l2(out, lin - b) < 2e-6
    # exact-erf GELU
    out = ops.gemm(a, w, b, gelu=True, out_f32=True, impl=impl)
    assert rel_l2(out, torch.nn.functional.gelu(lin)) < 5e-6
    # residual, in place on the fp32 stream
    x = res.clone()
    ops.gemm(a, w, b, out=x, residual=x, out_f32=True, impl=impl)
    assert rel_l2(x, lin + res) < 2e-6
    # broadcast residual row (aggregator CLS token)
    out = ops.gemm(a, w, b, residual=res[:1].contiguous(), out_f32=True, impl=impl)
    assert rel_l2(out, lin + res[:1]) < 2e-6
    # strided A (a column block of a wider activation) and strided output
    wide = _bf(torch.randn(M, 3 * K, device='cuda', generator=g))
    dst = torch.zeros(M, 2 * N, device='cuda', dtype=torch.bfloat16)
    ops.gemm(wide[:, K:2 * K], w, b, out=dst[:, N:], impl=impl)
    assert rel_l2(dst[:, N:].float(), wide[:, K:2 * K].float() @ w.float().T + b) < 4e-3
    assert float(dst[:, :N].abs().max()) == 0.0
    torch.cuda.synchronize()


@pytest.mark.parametrize('impl', [pytest.param(0, id='auto'), pytest.param(2, id='tc1cta'), pytest.param(3, id='tcpair')])
@pytest.mark.parametrize('M,N,gelu', [(1000, 2304, False), (517, 3072, True), (100, 1536, False), (1569 * 3, 768, False)])
def test_gemm_layernorm_fusion(cuda_device, impl, M, N, gelu):
    """sfb_gemm_bf16_ln: a residual GEMM that EMITs the bf16 copy + row statistics of its fp32 output, feeding a GEMM that applies the
    LayerNorm in its epilogue through folded weights (LN_FOLD) - against the definitions in the header, and against the unfused pair
    LayerNorm kernel + plain GEMM (vit_helper.py:366-375: x += proj(...); y = Linear(norm(x)))."""
    from synchformer_b200 import ops
    D = 768
    g = torch.Generator(device='cuda').manual_seed(M + N)
    att = _bf(torch.randn(M, D, device='cuda', generator=g))
    wp = _bf(torch.randn(D, D, device='cuda', generator=g) * 0.04)
    bp = torch.randn(D, device='cuda', generator=g) * 0.1
    x0 = torch.randn(M, D, device='cuda', generator=g) * 1.5 + 0.3                        # residual stream with a non-zero mean
    x0[:, 5] += 20.0                                                                       # and one outlier channel
    gamma, beta = torch.rand(D, device='cuda', generator=g) + 0.5, torch.randn(D, device='cuda', generator=g) * 0.2
    w0 = torch.randn(N, D, device='cuda', generator=g) * 0.03
    b0 = torch.randn(N, device='cuda', generator=g) * 0.1
    eps = 1e-6
    # producer: x = x0 + att wp^T + bp, plus bf16 copy and per-row (sum, sum of squares) per 64-column group
    xb, st = torch.empty(M, D, device='cuda', dtype=torch.bfloat16), torch.empty(M, D // 64, 2, device='cuda')
    x = ops.gemm(att, wp, bp, residual=x0, out_f32=True, emit_ln=(xb, st), impl=impl)
    x_plain = ops.gemm(att, wp, bp, residual=x0, out_f32=True, impl=impl)
    torch.cuda.synchronize()
    assert torch.equal(x, x_plain)                                                         # the fp32 result is unchanged by the emission
    assert torch.equal(xb, x.to(torch.bfloat16))                                           # bit-exact copy
    grp = x.double().view(M, D // 64, 64)
    ref_st = torch.stack([grp.sum(-1), (grp * grp).sum(-1)], dim=-1)
    assert (st.double() - ref_st).abs().max() <= 2e-5 * ref_st.abs().max()
    # consumer with folded weights
    wf = _bf(w0 * gamma.unsqueeze(0))
    cs = wf.float().sum(1).contiguous()
    bf = (b0 + w0 @ beta).contiguous()
    y = ops.gemm(xb, wf, bf, gelu=gelu, ln_fold=(st, cs, eps), impl=impl)
    torch.cuda.synchronize()
    mean = x.double().mean(-1, keepdim=True)
    rstd = torch.rsqrt(x.double().var(-1, unbiased=False, keepdim=True) + eps)
    ref = ((xb.double() - mean) * rstd) @ wf.double().T + bf.double()                      # the header's definition
    ref = torch.nn.functional.gelu(ref) if gelu else ref
    assert y.dtype == torch.bfloat16 and rel_l2(y.float(), ref) < 4e-3
    # the CUDA-core cross-check kernel implements the same LN_FOLD epilogue
    y1 = ops.gemm(xb, wf, bf, gelu=gelu, ln_fold=(st, cs, eps), impl=1)
    assert rel_l2(y1.float(), ref) < 4e-3
    # and the fused pair equals the unfused pair (LayerNorm kernel + plain GEMM on unfolded weights) to bf16 rounding
    ln = ops.layernorm(x, gamma, beta, eps)
    y_unfused = ops.gemm(ln, _bf(w0), b0, gelu=gelu, impl=impl)
    exact = torch.nn.functional.layer_norm(x.double(), (D,), gamma.double(), beta.double(), eps) @ w0.double().T + b0.double()
    exact = torch.nn.functional.gelu(exact) if gelu else exact
    e_f, e_u = rel_l2(y.float(), exact), rel_l2(y_unfused.float(), exact)
    assert e_f < 8e-3 and e_f < 1.5 * e_u + 1e-3, (e_f, e_u)
    # rows that no EMIT_LN epilogue produced: sfb_rowstats_cast (one partial per row)
    xb1, st1 = ops.rowstats_cast(x)
    assert torch.equal(xb1, xb)
    y2 = ops.gemm(xb1, wf, bf, gelu=gelu, ln_fold=(st1, cs, eps), impl=impl)
    assert rel_l2(y2.float(), ref) < 4e-3


@pytest.mark.parametrize('impl', [pytest.param(0, id='auto'), pytest.param(2, id='tc1cta'), pytest.param(3, id='tcpair')])
@pytest.mark.parametrize('M,N,K', [(1, 768, 768), (77, 40, 72), (300, 96, 256), (129, 1536, 768), (1000, 2304, 768), (257, 264, 64), (4099, 768, 3072)])
def test_gemm_fp32_epilogues_tails_and_strides(cuda_device, monkeypatch, impl, M, N, K):
    """fp32-output epilogues (TMA: residual box in / tile out; per-thread: SFB_GEMM_F32_TMA=0) on ragged M / N, with the residual stream
    updated in place, a residual with its own row stride, a strided output (a column block of a wider matrix, untouched outside), and the
    broadcast residual row that always takes the per-thread path - each against fp32 torch, and the two epilogues against each other."""
    from synchformer_b200 import ops
    g = torch.Generator(device='cuda').manual_seed(M * 13 + N)
    a = _bf(torch.randn(M, K, device='cuda', generator=g))
    w = _bf(torch.randn(N, K, device='cuda', generator=g) * 0.05)
    b = torch.randn(N, device='cuda', generator=g)
    res_wide = torch.randn(M, N + 8, device='cuda', generator=g)
    res = res_wide[:, :N]                                                   # row stride N + 8
    lin = a.float() @ w.float().T + b
    tol = 2.5e-6 * max(1.0, (K / 768) ** 0.5)                               # fp32 accumulation in a different order than torch's
    outs = {}
    for tma in ('1', '0'):
        monkeypatch.setenv('SFB_GEMM_F32_TMA', tma)
        plain = ops.gemm(a, w, b, out_f32=True, impl=impl)
        assert rel_l2(plain, lin) < tol
        with_res = ops.gemm(a, w, b, residual=res, out_f32=True, impl=impl)
        assert rel_l2(with_res, lin + res) < tol
        x = res.contiguous().clone()
        ops.gemm(a, w, b, out=x, residual=x, out_f32=True, impl=impl)          # in place on the stream
        assert rel_l2(x, lin + res) < tol
        wide = torch.full((M, N + 16), 7.0, device='cuda')
        ops.gemm(a, w, None, out=wide[:, 8:8 + N], residual=res, out_f32=True, gelu=True, impl=impl)
        assert rel_l2(wide[:, 8:8 + N], torch.nn.functional.gelu(lin - b) + res) < 2 * tol
        assert float((wide[:, :8] - 7.0).abs().max()) == 0.0 and float((wide[:, 8 + N:] - 7.0).abs().max()) == 0.0
        bc = ops.gemm(a, w, b, residual=res[:1].contiguous(), out_f32=True, impl=impl)
        assert rel_l2(bc, lin + res[:1]) < tol
        outs[tma] = (plain, with_res, x)
    torch.cuda.synchronize()
    for u, v in zip(outs['1'], outs['0']):
        assert torch.equal(u, v)                                              # same accumulator, same fp32 additions in the same order


def test_gemm_full_size_linearity_property(cuda_device):
    """At the benchmark's row count (64 clips x 8 segments x 1569 tokens) the result cannot be compared with a CPU oracle in
    seconds; use linearity instead: (A + A') W == A W + A' W up to fp32 accumulation order, on a strided sample of rows."""
    from synchformer_b200 import ops
    M, N, K = 64 * 8 * 1569, 768, 768
    g = torch.Generator(device='cuda').manual_seed(11)
    # small integers are exact in bf16 and their products/sums are exact in fp32 -> bit-exact property
    a1 = torch.randint(-4, 5, (M, K), device='cuda', generator=g).to(torch.bfloat16)
    a2 = torch.randint(-4, 5, (M, K), device='cuda', generator=g).to(torch.bfloat16)
    w = torch.randint(-3, 4, (N, K), device='cuda', generator=g).to(torch.bfloat16)
    y1 = ops.gemm(a1, w, None, out_f32=True)
    y2 = ops.gemm(a2, w, None, out_f32=True)
    y12 = ops.gemm((a1.float() + a2.float()).to(torch.bfloat16), w, None, out_f32=True)
    assert torch.equal(y12, y1 + y2)
    rows = torch.arange(0, M, 4099, device='cuda')
    assert torch.equal(y1[rows], a1[rows].float() @ w.float().T)
    assert torch.equal(y1[-3:], a1[-3:].float() @ w.float().T)        # the ragged last tile


# -------------------------------------------------------------------------------------------------- LayerNorm
@pytest.mark.parametrize('eps', [1e-6, 1e-12, 1e-5])
def test_layernorm(cuda_device, eps):
    from synchformer_b200 import ops
    g = torch.Generator(device='cuda').manual_seed(5)
    x = torch.randn(1003, 768, device='cuda', generator=g) * 3 + 0.5
    w = torch.randn(768, device='cuda', generator=g)
    b = torch.randn(768, device='cuda', generator=g)
    ref = torch.nn.functional.layer_norm(x, (768,), w, b, eps)
    out = ops.layernorm(x, w, b, eps, out_f32=True)
    assert (out - ref).abs().max() < 2e-5
    out = ops.layernorm(x, w, b, eps)
    assert out.dtype == torch.bfloat16 and rel_l2(out.float(), ref) < 4e-3


def test_layernorm_gather_and_double(cuda_device):
    from synchformer_b200 import ops
    g = torch.Generator(device='cuda').manual_seed(6)
    n = 3
    x = torch.randn(n * 1569, 768, device='cuda', generator=g)
    w1, b1, w2, b2 = [torch.randn(768, device='cuda', generator=g) for _ in range(4)]
    out = ops.layernorm(x, w1, b1, 1e-6, rows=n * 1568, group=1568, group_stride=1569, offset=1, gamma2=w2, beta2=b2, eps2=1e-6, out_f32=True)
    src = x.view(n, 1569, 768)[:, 1:].reshape(-1, 768)
    ref = torch.nn.functional.layer_norm(torch.nn.functional.layer_norm(src, (768,), w1, b1, 1e-6), (768,), w2, b2, 1e-6)
    assert out.shape == (n * 1568, 768) and (out - ref).abs().max() < 5e-5


# -------------------------------------------------------------------------------------------------- attention
def _ref_attention(q, k, v, scale):
    s = (q.float() @ k.float().transpose(-1, -2)) * scale
    return torch.softmax(s, -1) @ v.float()


@pytest.mark.parametrize('impl', [pytest.param(0, id='tc'), pytest.param(1, id='simple')])
@pytest.mark.parametrize('mode', ['time', 'space', 'cls'])
def test_divided_attention_views(cuda_device, impl, mode):
    """Motionformer divided attention on the fused (n*1569, 2304) qkv layout (vit_helper.py:100-158)."""
    from synchformer_b200 import ops
    n, D = 2, 768
    g = torch.Generator(device='cuda').manual_seed(17)
    qkv = _bf(torch.randn(n * 1569, 3 * D, device='cuda', generator=g))
    att = torch.zeros(n * 1569, D, device='cuda', dtype=torch.bfloat16)
    row, seg = 3 * D, 1569 * 3 * D
    t = qkv.float().view(n, 1569, 3, 12, 64)
    q, k, v = t[:, :, 0].permute(0, 2, 1, 3), t[:, :, 1].permute(0, 2, 1, 3), t[:, :, 2].permute(0, 2, 1, 3)   # (n, h, 1569, 64)
    if mode == 'cls':
        ops.attention(qkv, qkv[:, D:], qkv[:, 2 * D:], att, q_strides=(seg, 0, row), kv_strides=(seg, 0, row), o_strides=(1569 * D, 0, D),
                      n_outer=n, n_inner=1, n_heads=12, head_dim=64, Lq=1, Lk=1569, scale=0.125, impl=impl)
        ref = _ref_attention(q[:, :, :1], k, v, 0.125)                                    # (n, h, 1, 64)
        got = att.view(n, 1569, 12, 64)[:, :1].permute(0, 2, 1, 3).float()
    else:
        q_, k_, v_ = [x[:, :, 1:].reshape(n, 12, 8, 196, 64) for x in (q, k, v)]
        ck, cv = k[:, :, :1], v[:, :, :1]
        if mode == 'time':
            q_, k_, v_ = [x.permute(0, 1, 3, 2, 4) for x in (q_, k_, v_)]                 # (n, h, 196, 8, 64)
            strides = dict(q_strides=(seg, row, 196 * row), kv_strides=(seg, row, 196 * row), o_strides=(1569 * D, D, 196 * D), n_inner=196, Lq=8, Lk=8)
        else:
            strides = dict(q_strides=(seg, 196 * row, row), kv_strides=(seg, 196 * row, row), o_strides=(1569 * D, 196 * D, D), n_inner=8, Lq=196, Lk=196)
        G = q_.shape[2]
        kk = torch.cat([ck.unsqueeze(2).expand(n, 12, G, 1, 64), k_], 3)
        vv = torch.cat([cv.unsqueeze(2).expand(n, 12, G, 1, 64), v_], 3)
        ref = _ref_attention(q_, kk, vv, 0.125)
        if mode == 'time':
            ref = ref.permute(0, 1, 3, 2, 4)
        ref = ref.reshape(n, 12, 1568, 64)
        ops.attention(qkv[1:], qkv[1:, D:], qkv[1:, 2 * D:], att[1:], n_outer=n, n_heads=12, head_dim=64, scale=0.125,
                      k_prefix=qkv[:, D:], v_prefix=qkv[:, 2 * D:], prefix_outer=seg, impl=impl, **strides)
        got = att.view(n, 1569, 12, 64)[:, 1:].permute(0, 2, 1, 3).float()
        assert float(att.view(n, 1569, D)[:, 0].abs().max()) == 0.0                       # CLS rows untouched by this call
    torch.cuda.synchronize()
    assert rel_l2(got, ref) < 6e-3, mode
    assert (got - ref).abs().max() < 3e-2


@pytest.mark.parametrize('q_gain,k_gain', [(9.0, 9.0), (0.01, 0.01), (30.0, 0.02), (1.0, 1.0)])
def test_space_attention_softmax_reference_value(cuda_device, q_gain, k_gain):
    """Softmax range handling of the tcgen05 space-attention kernel against fp32 torch: huge logits (gains 9 x 9: near one-hot rows, scores
    of several hundred), tiny and mixed magnitudes, and rows whose queries differ by orders of magnitude inside one warp - nothing may
    overflow, underflow to a zero row sum, or lose the bf16 precision of the probabilities."""
    from synchformer_b200 import ops
    n, D = 1, 768
    g = torch.Generator(device='cuda').manual_seed(int(q_gain * 100 + k_gain * 7))
    raw = torch.randn(n * 1569, 3 * D, device='cuda', generator=g)
    raw[:, :D] *= q_gain
    raw[:, D:2 * D] *= k_gain
    raw[5::7, :D] *= 40.0                                                          # a few rows with much larger queries
    qkv = _bf(raw)
    att = torch.zeros(n * 1569, D, device='cuda', dtype=torch.bfloat16)
    row, seg = 3 * D, 1569 * 3 * D
    t = qkv.float().view(n, 1569, 3, 12, 64)
    q, k, v = t[:, :, 0].permute(0, 2, 1, 3), t[:, :, 1].permute(0, 2, 1, 3), t[:, :, 2].permute(0, 2, 1, 3)
    q_, k_, v_ = [x[:, :, 1:].reshape(n, 12, 8, 196, 64) for x in (q, k, v)]
    kk = torch.cat([k[:, :, :1].unsqueeze(2).expand(n, 12, 8, 1, 64), k_], 3)
    vv = torch.cat([v[:, :, :1].unsqueeze(2).expand(n, 12, 8, 1, 64), v_], 3)
    ref = _ref_attention(q_, kk, vv, 0.125).reshape(n, 12, 1568, 64)
    ops.attention(qkv[1:], qkv[1:, D:], qkv[1:, 2 * D:], att[1:], q_strides=(seg, 196 * row, row), kv_strides=(seg, 196 * row, row),
                  o_strides=(1569 * D, 196 * D, D), n_outer=n, n_inner=8, n_heads=12, head_dim=64, Lq=196, Lk=196, scale=0.125,
                  k_prefix=qkv[:, D:], v_prefix=qkv[:, 2 * D:], prefix_outer=seg)
    torch.cuda.synchronize()
    got = att.view(n, 1569, 12, 64)[:, 1:].permute(0, 2, 1, 3).float()
    assert torch.isfinite(got).all()
    assert rel_l2(got, ref) < 8e-3, (q_gain, k_gain, rel_l2(got, ref))


@pytest.mark.parametrize('direction', [1.0, -1.0])
def test_space_attention_single_pass_softmax_raises_its_reference(cuda_device, direction):
    """The tcgen05 space-attention kernel exponentiates against a reference value taken from the first 32 keys and raises it when a later
    chunk of keys exceeds it by more than 2^64 (already stored probabilities are rescaled by an exact power of two).  Keys whose scores
    grow with the key index (every chunk, the 16-key tail and the CLS prefix key trigger a raise) or fall with it (no raise at all), with a
    step of ~90 log2 units per 32 keys, against fp32 torch."""
    from synchformer_b200 import ops
    n, D = 1, 768
    g = torch.Generator(device='cuda').manual_seed(11)
    raw = torch.randn(n * 1569, 3 * D, device='cuda', generator=g)
    u = torch.randn(12, 64, device='cuda', generator=g)
    u = u / u.norm(dim=1, keepdim=True)
    # q = a u + noise, k_j = ramp(j) u + noise: score ~ a ramp(j) / 8; the ramp also differs per frame and is largest for the CLS row
    pos = torch.arange(1569, device='cuda', dtype=torch.float32)
    ramp = direction * ((pos - 1) % 196) * 1.0
    ramp[0] = direction * 230.0
    raw[:, :D] = raw[:, :D] + 125.0 * u.reshape(1, D)
    raw[:, D:2 * D] = raw[:, D:2 * D] * 0.5 + ramp[:, None] * u.reshape(1, D) * 0.125
    qkv = _bf(raw)
    att = torch.zeros(n * 1569, D, device='cuda', dtype=torch.bfloat16)
    row, seg = 3 * D, 1569 * 3 * D
    t = qkv.float().view(n, 1569, 3, 12, 64)
    q, k, v = t[:, :, 0].permute(0, 2, 1, 3), t[:, :, 1].permute(0, 2, 1, 3), t[:, :, 2].permute(0, 2, 1, 3)
    q_, k_, v_ = [x[:, :, 1:].reshape(n, 12, 8, 196, 64) for x in (q, k, v)]
    kk = torch.cat([k[:, :, :1].unsqueeze(2).expand(n, 12, 8, 1, 64), k_], 3)
    vv = torch.cat([v[:, :, :1].unsqueeze(2).expand(n, 12, 8, 1, 64), v_], 3)
    logits = torch.einsum('bhfqd,bhfkd->bhfqk', q_, kk) * 0.125 * 1.4426950408889634
    spread = (logits[..., 33:].amax(-1) - logits[..., 1:33].amax(-1))             # keys 0..31 of the kernel's order are columns 1..32 here
    if direction > 0:
        assert spread.min() > 64.0 * 2, spread.min()                             # every row has to raise its reference, several times
    else:
        assert spread.max() < 0.0
    ref = _ref_attention(q_, kk, vv, 0.125).reshape(n, 12, 1568, 64)
    ops.attention(qkv[1:], qkv[1:, D:], qkv[1:, 2 * D:], att[1:], q_strides=(seg, 196 * row, row), kv_strides=(seg, 196 * row, row),
                  o_strides=(1569 * D, 196 * D, D), n_outer=n, n_inner=8, n_heads=12, head_dim=64, Lq=196, Lk=196, scale=0.125,
                  k_prefix=qkv[:, D:], v_prefix=qkv[:, 2 * D:], prefix_outer=seg)
    torch.cuda.synchronize()
    got = att.view(n, 1569, 12, 64)[:, 1:].permute(0, 2, 1, 3).float()
    assert torch.isfinite(got).all()
    assert rel_l2(got, ref) < 8e-3, rel_l2(got, ref)


@pytest.mark.parametrize('impl', [pytest.param(0, id='tc'), pytest.param(1, id='simple')])
@pytest.mark.parametrize('L,heads,hd', [(74, 12, 64), (198, 8, 96), (114, 8, 96), (30, 8, 96), (16, 12, 64), (17, 12, 64)])
def test_plain_self_attention(cuda_device, impl, L, heads, hd):
    """AST (74 x 74, hd 64, modeling_ast.py:145-184) and sync (T x T, hd 96, modules/transformer.py:58-76) layouts."""
    from synchformer_b200 import ops
    B, D = 3, 768
    g = torch.Generator(device='cuda').manual_seed(L)
    qkv = _bf(torch.randn(B * L, 3 * D, device='cuda', generator=g) * 1.5)
    att = torch.empty(B * L, D, device='cuda', dtype=torch.bfloat16)
    scale = 1.0 / math.sqrt(hd)
    ops.attention(qkv, qkv[:, D:], qkv[:, 2 * D:], att, q_strides=(L * 3 * D, 0, 3 * D), kv_strides=(L * 3 * D, 0, 3 * D), o_strides=(L * D, 0, D),
                  n_outer=B, n_inner=1, n_heads=heads, head_dim=hd, Lq=L, Lk=L, scale=scale, impl=impl)
    t = qkv.float().view(B, L, 3, heads, hd)
    ref = _ref_attention(t[:, :, 0].permute(0, 2, 1, 3), t[:, :, 1].permute(0, 2, 1, 3), t[:, :, 2].permute(0, 2, 1, 3), scale)
    got = att.view(B, L, heads, hd).permute(0, 2, 1, 3).float()
    torch.cuda.synchronize()
    assert rel_l2(got, ref) < 6e-3
    assert (got - ref).abs().max() < 3e-2


@pytest.mark.parametrize('impl', [pytest.param(0, id='tc'), pytest.param(1, id='simple')])
@pytest.mark.parametrize('Lk,inner,row_rows', [(196, 8, 1), (12, 6, 6)])
def test_aggregator_cls_query_attention(cuda_device, Lk, inner, row_rows, impl):
    """One shared CLS query per head against strided K/V groups plus a prefix CLS key/value (motionformer.py:301-334)."""
    from synchformer_b200 import ops
    n, D = 3, 768
    rows_per_seg = Lk * inner
    g = torch.Generator(device='cuda').manual_seed(Lk)
    kv = _bf(torch.randn(n * rows_per_seg, 2 * D, device='cuda', generator=g))
    cls_qkv = _bf(torch.randn(1, 3 * D, device='cuda', generator=g))
    out = torch.empty(n * inner, D, device='cuda', dtype=torch.bfloat16)
    inner_rows = Lk if row_rows == 1 else 1
    ops.attention(cls_qkv, kv, kv[:, D:], out, q_strides=(0, 0, 0), kv_strides=(rows_per_seg * 2 * D, inner_rows * 2 * D, row_rows * 2 * D),
                  o_strides=(inner * D, D, D), n_outer=n, n_inner=inner, n_heads=12, head_dim=64, Lq=1, Lk=Lk, scale=0.125,
                  k_prefix=cls_qkv[:, D:], v_prefix=cls_qkv[:, 2 * D:], prefix_outer=0, impl=impl)
    kvf = kv.float().view(n, rows_per_seg, 2, 12, 64)
    if row_rows == 1:
        grp = kvf.view(n, inner, Lk, 2, 12, 64)
    else:
        grp = kvf.view(n, Lk, inner, 2, 12, 64).permute(0, 2, 1, 3, 4, 5)
    k = torch.cat([cls_qkv.float()[:, D:2 * D].view(1, 1, 1, 12, 64).expand(n, inner, 1, 12, 64), grp[:, :, :, 0]], 2).permute(0, 1, 3, 2, 4)
    v = torch.cat([cls_qkv.float()[:, 2 * D:].view(1, 1, 1, 12, 64).expand(n, inner, 1, 12, 64), grp[:, :, :, 1]], 2).permute(0, 1, 3, 2, 4)
    q = cls_qkv.float()[:, :D].view(1, 1, 12, 1, 64).expand(n, inner, 12, 1, 64)
    ref = _ref_attention(q, k, v, 0.125).reshape(n * inner, D)
    torch.cuda.synchronize()
    assert rel_l2(out.float(), ref) < 6e-3


# ------------------------------------------------------------------------------------- embeddings and tokens
@pytest.mark.parametrize('dtype', [torch.float32, torch.float16, torch.bfloat16, torch.uint8])
def test_im2col_video(cuda_device, dtype):
    from synchformer_b200 import ops
    n = 2
    g = torch.Generator(device='cuda').manual_seed(1)
    if dtype == torch.uint8:
        vis = torch.randint(0, 256, (n, 16, 3, 224, 224), device='cuda', generator=g, dtype=torch.uint8)
        visf = ((vis.double() / 255.0 - 0.5) / 0.5).float()           # dataset/transforms.py:647-669, exact value rounded once
    else:
        vis = (torch.rand(n, 16, 3, 224, 224, device='cuda', generator=g) * 2 - 1).to(dtype)
        visf = vis.float()
    a = ops.im2col_video(vis)
    ref = visf.view(n, 8, 2, 3, 14, 16, 14, 16).permute(0, 1, 4, 6, 3, 2, 5, 7).reshape(n * 1568, 1536)
    torch.cuda.synchronize()
    if dtype == torch.uint8:   # the fused normalisation rounds once in fp32: at most 1 bf16 ulp away, and only on rounding ties
        diff = (a.float() - ref.to(torch.bfloat16).float()).abs()
        assert float(diff.max()) <= 2.0 ** -8 and float((diff > 0).float().mean()) < 1e-3
    else:
        assert torch.equal(a, ref.to(torch.bfloat16))


def test_patch_embed_equals_conv3d(cuda_device):
    """im2col + GEMM + token assembly == Conv3d(k=s=(2,16,16)) + cls + separate pos-emb (vit_helper.py:436-445, video_model_builder.py:221-254)."""
    from synchformer_b200 import ops
    n = 2
    g = torch.Generator(device='cuda').manual_seed(2)
    vis = _bf(torch.rand(n, 16, 3, 224, 224, device='cuda', generator=g) * 2 - 1)
    w = _bf(torch.randn(768, 3, 2, 16, 16, device='cuda', generator=g) * 0.03)
    b, cls = torch.randn(768, device='cuda', generator=g), torch.randn(1, 1, 768, device='cuda', generator=g)
    pos, tmp = torch.randn(1, 197, 768, device='cuda', generator=g), torch.randn(1, 8, 768, device='cuda', generator=g)
    patch = ops.gemm(ops.im2col_video(vis), w.view(768, -1).contiguous(), b, out_f32=True)
    x = ops.video_tokens(patch, cls, pos, tmp, n).view(n, 1569, 768)
    conv = torch.nn.functional.conv3d(vis.float().permute(0, 2, 1, 3, 4), w.float(), b, stride=(2, 16, 16)).flatten(2).transpose(1, 2)
    ref = torch.cat([cls.expand(n, 1, 768), conv], 1)
    total = torch.cat([pos[:, :1], pos[:, 1:].repeat(1, 8, 1) + tmp.repeat_interleave(196, 1)], 1)
    ref = ref + total
    torch.cuda.synchronize()
    assert rel_l2(x, ref) < 1e-5


def test_ast_patch_embed_equals_conv2d(cuda_device):
    from synchformer_b200 import ops
    n = 3
    g = torch.Generator(device='cuda').manual_seed(3)
    spec = torch.randn(n, 128, 66, device='cuda', generator=g)
    w = _bf(torch.randn(768, 1, 16, 16, device='cuda', generator=g) * 0.05)
    b = torch.randn(768, device='cuda', generator=g)
    cls, dist = torch.randn(1, 1, 768, device='cuda', generator=g), torch.randn(1, 1, 768, device='cuda', generator=g)
    pos = torch.randn(1, 74, 768, device='cuda', generator=g)
    a = ops.im2col_ast(spec)
    patch = ops.gemm(a, w.view(768, 256).contiguous(), b, out_f32=True)
    x = ops.ast_tokens(patch, cls, dist, pos, n).view(n, 74, 768)
    conv = torch.nn.functional.conv2d(_bf(spec).float().unsqueeze(1), w.float(), b, stride=(10, 10)).flatten(2).transpose(1, 2)   # (n, 72, 768)
    ref = torch.cat([cls.expand(n, 1, 768), dist.expand(n, 1, 768), conv], 1) + pos
    torch.cuda.synchronize()
    assert rel_l2(x, ref) < 1e-5


def test_sync_tokens_and_head(cuda_device):
    from synchformer_b200 import ops
    B, S = 3, 4
    T = 2 + 14 * S
    g = torch.Generator(device='cuda').manual_seed(4)
    r = lambda *s: torch.randn(*s, device='cuda', generator=g)
    v, a = r(B, 8 * S, 768), r(B, 6 * S, 768)
    vw, vb, aw, ab, off, mod, pos = r(768), r(768), r(768), r(768), r(1, 1, 768), r(1, 1, 768), r(1, T, 768)
    x = ops.sync_tokens(v, a, vw, vb, aw, ab, 1e-5, off, mod, pos, B, S).view(B, T, 768)
    ln = torch.nn.functional.layer_norm
    ref = torch.cat([off.expand(B, 1, 768), ln(v, (768,), vw, vb, 1e-5), mod.expand(B, 1, 768), ln(a, (768,), aw, ab, 1e-5)], 1) + pos
    assert (x - ref).abs().max() < 3e-5
    lw, lb, W, bb = r(768), r(768), r(21, 768) * 0.05, r(21)
    logits = ops.sync_head(x.view(B * T, 768), T, lw, lb, 1e-5, W, bb, B)
    ref_logits = ln(ref[:, 0], (768,), lw, lb, 1e-5) @ W.T + bb
    torch.cuda.synchronize()
    assert (logits - ref_logits).abs().max() < 1e-4


def test_cast(cuda_device):
    from synchformer_b200 import ops
    x = torch.randn(1000, 768, device='cuda')
    assert torch.equal(ops.cast_bf16(x), x.to(torch.bfloat16))


# ------------------------------------------------------------------------------------------------ mel front-end
def test_mel_frontend_against_oracle_and_reference_golden(cuda_device, golden_dir):
    import os
    from oracle import synchformer_oracle as O
    from synchformer_b200 import ops, synth
    wave = synth.synthetic_waveform(2, 2, 0)
    mel = ops.mel_frontend(wave.cuda()).cpu()
    assert mel.shape == (2, 2, 128, 66)
    oracle = O.mel_frontend(wave).float()
    golden = torch.from_numpy(np.load(os.path.join(golden_dir, 'mel_b2s2.npz'))['mel'])
    # tolerance: the normalised log-mel spans about [-1.1, 1.3]; torchaudio's own fp32 FFT sits 2.3e-4 from the fp64 oracle
    assert (mel - oracle).abs().max() < 5e-4
    assert (mel - golden).abs().max() < 1e-3
    # random (broadband) input exercises every bin and the reflect padding at both ends
    g = torch.Generator().manual_seed(9)
    noise = torch.randn(5, 10240, generator=g) * 0.3
    assert (ops.mel_frontend(noise.cuda()).cpu() - O.mel_frontend(noise).float()).abs().max() < 5e-4


def test_space_attention_with_fused_cls_query(cuda_device):
    """The CLS query (vit_helper.py:124: one query against all 1569 keys of its segment) rides along with the 8 per-frame space-attention
    problems of the tcgen05 kernel and is merged from their partial softmax states; result must equal the stand-alone computation."""
    from synchformer_b200 import ops
    n, D = 3, 768
    g = torch.Generator(device='cuda').manual_seed(23)
    qkv = _bf(torch.randn(n * 1569, 3 * D, device='cuda', generator=g))
    att = torch.zeros(n * 1569, D, device='cuda', dtype=torch.bfloat16)
    row, seg = 3 * D, 1569 * 3 * D
    fused = ops.attention(qkv[1:], qkv[1:, D:], qkv[1:, 2 * D:], att[1:], q_strides=(seg, 196 * row, row), kv_strides=(seg, 196 * row, row),
                          o_strides=(1569 * D, 196 * D, D), n_outer=n, n_inner=8, n_heads=12, head_dim=64, Lq=196, Lk=196, scale=0.125,
                          k_prefix=qkv[:, D:], v_prefix=qkv[:, 2 * D:], prefix_outer=seg,
                          q_extra=qkv, q_extra_outer=seg, extra_out=att, extra_out_outer=1569 * D)
    assert fused, 'the tcgen05 space-attention kernel should take the extra query'
    t = qkv.float().view(n, 1569, 3, 12, 64)
    q, k, v = t[:, :, 0].permute(0, 2, 1, 3), t[:, :, 1].permute(0, 2, 1, 3), t[:, :, 2].permute(0, 2, 1, 3)
    ref_cls = _ref_attention(q[:, :, :1], k, v, 0.125)                       # (n, h, 1, 64)
    got_cls = att.view(n, 1569, 12, 64)[:, :1].permute(0, 2, 1, 3).float()
    torch.cuda.synchronize()
    assert rel_l2(got_cls, ref_cls) < 6e-3
    # and the regular rows are unchanged by the extra row
    att2 = torch.zeros_like(att)
    ops.attention(qkv[1:], qkv[1:, D:], qkv[1:, 2 * D:], att2[1:], q_strides=(seg, 196 * row, row), kv_strides=(seg, 196 * row, row),
                  o_strides=(1569 * D, 196 * D, D), n_outer=n, n_inner=8, n_heads=12, head_dim=64, Lq=196, Lk=196, scale=0.125,
                  k_prefix=qkv[:, D:], v_prefix=qkv[:, 2 * D:], prefix_outer=seg)
    assert torch.equal(att.view(n, 1569, D)[:, 1:], att2.view(n, 1569, D)[:, 1:])
