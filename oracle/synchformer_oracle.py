"""CPU oracle for the Synchformer segment-batched audio-visual forward path.

TEST INFRASTRUCTURE, NOT PRODUCT.  Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import this file; the shipped path (`synchformer_b200/`) never does.

It is an independent, functional restatement (plain `torch` CPU tensor algebra, fp32 or fp64, no nn.Module,
no reference code) of what the reference computes for `model.sync_model.Synchformer.forward` in eval mode.
Every function cites the reference lines it follows (paths relative to the reference root).

Parity pin: the reference ships no tests / golden vectors for this path, so the oracle is
pinned against OUTPUTS OF THE REFERENCE ITSELF, generated in the build container by
`tests/golden/make_golden.py` (imports the unmodified reference from /root/reference) and committed as
`tests/golden/*.npz`; `tests/test_oracle_golden.py` checks the oracle against them on every CPU run, and
`tests/test_oracle_vs_reference.py` re-runs the live reference when /root/reference is present.

Weights come in as the reference's own `state_dict` naming, i.e. a mapping
`name -> tensor`; the oracle never instantiates modules.
"""
import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

# ----------------------------------------------------------------------------------------------------------
# fixed hyper-parameters of the path (configs/sync.yaml:3-59, motionformer_src/divided_224_16x4.yaml:47-64,
# transformers ASTConfig defaults)
# ----------------------------------------------------------------------------------------------------------
D = 768
V_HEADS, V_DEPTH, V_FRAMES, V_SPACE = 12, 12, 8, 196       # 8 temporal tokens x 14x14 spatial tokens
A_HEADS, A_DEPTH, A_F, A_T = 12, 12, 12, 6                 # 12 freq x 6 time patches of a 128x66 mel
S_HEADS, S_DEPTH = 8, 3
EPS_V, EPS_A, EPS_S = 1e-6, 1e-12, 1e-5                    # video_model_builder.py:39 | ASTConfig | nn.LayerNorm default
MEL = dict(sr=16000, n_fft=1024, win=400, hop=160, n_mels=128, f_min=0.0, f_max=8000.0,
           log_eps=1e-6, max_t=66, mean=-4.2677393, std=4.5689974)   # configs/sync.yaml:183-197


def _ln(x: Tensor, w: Tensor, b: Tensor, eps: float) -> Tensor:
    """y = (x - mean) / sqrt(biased_var + eps) * w + b over the last dim."""
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def _lin(x: Tensor, w: Tensor, b: Optional[Tensor]) -> Tensor:
    """y = x @ w^T + b."""
    return F.linear(x, w, b)


def _gelu(x: Tensor) -> Tensor:
    """exact erf GELU 0.5 x (1 + erf(x / sqrt 2)): nn.GELU() / HF GELUActivation."""
    return F.gelu(x)


def _softmax_av(q: Tensor, k: Tensor, v: Tensor) -> Tensor:
    """(..., Lq, d), (..., Lk, d), (..., Lk, d) -> (..., Lq, d); scale already applied by the caller."""
    return torch.softmax(q @ k.transpose(-1, -2), dim=-1) @ v


def _cast_sd(sd: Dict[str, Tensor], dtype, device='cpu') -> Dict[str, Tensor]:
    return {k: v.detach().to(device, dtype) for k, v in sd.items()}


# ----------------------------------------------------------------------------------------------------------
# a1: mel front-end.  dataset/transforms.py:815-871 (AudioMelSpectrogram -> AudioLog -> PadOrTruncate ->
# AudioNormalizeAST) with the parameters of configs/sync.yaml:183-197.  The arithmetic lives in torchaudio
# (not vendored; pinned torchaudio==2.0.0, conda_env.yml:226): transforms.MelSpectrogram defaults = centered
# STFT with reflect padding n_fft//2, periodic Hann(win_length) zero-padded symmetrically to n_fft, power 2,
# HTK mel scale, norm=None, f_min 0, f_max sr/2, filterbank (n_fft//2+1, n_mels) of triangles.
# ----------------------------------------------------------------------------------------------------------
def mel_filterbank(dtype=torch.float64) -> Tensor:
    """torchaudio.functional.melscale_fbanks(513, 0, 8000, 128, 16000, norm=None, mel_scale='htk') -> (513, 128)."""
    n_freqs = MEL['n_fft'] // 2 + 1
    all_freqs = torch.linspace(0, MEL['sr'] // 2, n_freqs, dtype=dtype)
    hz2mel = lambda f: 2595.0 * math.log10(1.0 + f / 700.0)
    m_pts = torch.linspace(hz2mel(MEL['f_min']), hz2mel(MEL['f_max']), MEL['n_mels'] + 2, dtype=dtype)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)            # (n_freqs, n_mels+2)
    down = -slopes[:, :-2] / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.clamp(torch.minimum(down, up), min=0.0)


def mel_window(dtype=torch.float64) -> Tensor:
    """periodic Hann(400) centred in a 1024 frame (torch.stft pads the window on both sides)."""
    n = torch.arange(MEL['win'], dtype=dtype)
    hann = 0.5 - 0.5 * torch.cos(2.0 * math.pi * n / MEL['win'])
    w = torch.zeros(MEL['n_fft'], dtype=dtype)
    left = (MEL['n_fft'] - MEL['win']) // 2
    w[left:left + MEL['win']] = hann
    return w


def mel_frontend(wave: Tensor, dtype=torch.float64) -> Tensor:
    """wave (..., 10240) -> normalised log-mel (..., 128, 66).  Direct DFT (no FFT library) so the restatement
    is self-contained; fp64 by default, so it is *more* exact than torchaudio's fp32 FFT."""
    x = wave.to('cpu', dtype)
    lead = x.shape[:-1]
    x = x.reshape(-1, x.shape[-1])
    n_fft, hop = MEL['n_fft'], MEL['hop']
    pad = n_fft // 2
    # reflect padding (torch.stft center=True, pad_mode='reflect')
    xp = torch.cat([x[:, 1:pad + 1].flip(-1), x, x[:, -pad - 1:-1].flip(-1)], dim=-1)
    n_frames = 1 + (xp.shape[-1] - n_fft) // hop
    frames = xp.unfold(-1, n_fft, hop)[:, :n_frames] * mel_window(dtype)       # (N, 65, 1024)
    k = torch.arange(n_fft // 2 + 1, dtype=dtype)
    n = torch.arange(n_fft, dtype=dtype)
    ang = 2.0 * math.pi * torch.outer(n, k) / n_fft                               # (1024, 513)
    re = frames @ torch.cos(ang)
    im = frames @ torch.sin(ang)
    power = re * re + im * im                                                     # (N, 65, 513)
    mel = (power @ mel_filterbank(dtype)).transpose(-1, -2)                        # (N, 128, 65)
    logmel = torch.log(mel + MEL['log_eps'])                                      # transforms.py:826-834
    if logmel.shape[-1] < MEL['max_t']:                                           # transforms.py:845-852 (pad value 0.0)
        logmel = F.pad(logmel, (0, MEL['max_t'] - logmel.shape[-1]), 'constant', 0.0)
    else:
        logmel = logmel[..., :MEL['max_t']]
    out = (logmel - MEL['mean']) / (2.0 * MEL['std'])                             # transforms.py:868-871
    return out.reshape(*lead, MEL['n_mels'], MEL['max_t'])


# ----------------------------------------------------------------------------------------------------------
# a3-a8: visual stream
# ----------------------------------------------------------------------------------------------------------
def video_patch_embed(sd, vis: Tensor) -> Tensor:
    """vis (BS, 16, 3, 224, 224) [frames, channels] -> tokens (BS, 1569, 768).
    vit_helper.py:436-445 (Conv3d k=s=(2,16,16) over (C,T,H,W)) + video_model_builder.py:221-254
    (prepend cls, 'separate' pos-emb: spatial table tiled over frames + temporal table repeat-interleaved)."""
    p = 'vfeat_extractor.'
    BS = vis.shape[0]
    w = sd[p + 'patch_embed_3d.proj.weight'].reshape(D, -1)                     # (768, 3*2*16*16), order (c, dt, dy, dx)
    x = vis.reshape(BS, 8, 2, 3, 14, 16, 14, 16)                                 # (bs, f, dt, c, y, dy, x, dx)
    x = x.permute(0, 1, 4, 6, 3, 2, 5, 7).reshape(BS, 8 * 196, 3 * 2 * 16 * 16)  # token order f*196 + y*14 + x
    tok = _lin(x, w, sd[p + 'patch_embed_3d.proj.bias'])
    pos = sd[p + 'pos_embed'][0]                                                 # (197, 768)
    tmp = sd[p + 'temp_embed'][0]                                                # (8, 768)
    total = pos[1:].unsqueeze(0) + tmp.unsqueeze(1)                              # (8, 196, 768)
    tok = tok + total.reshape(1, 8 * 196, D)
    cls = (sd[p + 'cls_token'][0] + pos[0:1]).unsqueeze(0).expand(BS, 1, D)
    return torch.cat([cls, tok], dim=1)


def divided_attention(sd, prefix: str, x: Tensor, mode: str) -> Tensor:
    """DividedAttention.forward, vit_helper.py:100-158, with einops_to = '(b n) f d' (time) or '(b f) n d' (space)."""
    BS, N, _ = x.shape
    h, d = V_HEADS, D // V_HEADS
    qkv = _lin(x, sd[prefix + 'qkv.weight'], sd[prefix + 'qkv.bias'])
    q, k, v = [t.reshape(BS, N, h, d).permute(0, 2, 1, 3) for t in qkv.chunk(3, dim=-1)]   # (BS, h, N, d)
    q = q * d ** -0.5                                                            # :113
    cls_out = _softmax_av(q[:, :, 0:1], k, v)                                    # :124  CLS attends to everything
    q_, k_, v_ = q[:, :, 1:], k[:, :, 1:], v[:, :, 1:]
    if mode == 'time':       # sequences of the 8 frames at one spatial location
        re = lambda t: t.reshape(BS, h, V_FRAMES, V_SPACE, d).permute(0, 1, 3, 2, 4)       # (BS, h, n, f, d)
    else:                    # sequences of the 196 locations of one frame
        re = lambda t: t.reshape(BS, h, V_FRAMES, V_SPACE, d)                               # (BS, h, f, n, d)
    q_, k_, v_ = re(q_), re(k_), re(v_)
    G = q_.shape[2]
    ck = k[:, :, 0:1].unsqueeze(2).expand(BS, h, G, 1, d)                        # :129-134 CLS key/value prepended
    cv = v[:, :, 0:1].unsqueeze(2).expand(BS, h, G, 1, d)
    out = _softmax_av(q_, torch.cat([ck, k_], dim=3), torch.cat([cv, v_], dim=3))
    if mode == 'time':
        out = out.permute(0, 1, 3, 2, 4)
    out = out.reshape(BS, h, V_FRAMES * V_SPACE, d)
    out = torch.cat([cls_out, out], dim=2).permute(0, 2, 1, 3).reshape(BS, N, D)  # :150-153
    return _lin(out, sd[prefix + 'proj.weight'], sd[prefix + 'proj.bias'])


def video_block(sd, i: int, x: Tensor, drop_path=None) -> Tensor:
    """DividedSpaceTimeBlock.forward, vit_helper.py:364-376: time (norm3) -> space (norm1) -> MLP (norm2).
    drop_path = (space multiplier, mlp multiplier), each (BS, 1, 1): timm DropPath in train mode (:371, :375); None = eval."""
    p = f'vfeat_extractor.blocks.{i}.'
    ms, mm = (1.0, 1.0) if drop_path is None else drop_path
    x = x + divided_attention(sd, p + 'timeattn.', _ln(x, sd[p + 'norm3.weight'], sd[p + 'norm3.bias'], EPS_V), 'time')
    x = x + ms * divided_attention(sd, p + 'attn.', _ln(x, sd[p + 'norm1.weight'], sd[p + 'norm1.bias'], EPS_V), 'space')
    hdn = _gelu(_lin(_ln(x, sd[p + 'norm2.weight'], sd[p + 'norm2.bias'], EPS_V), sd[p + 'mlp.fc1.weight'], sd[p + 'mlp.fc1.bias']))
    return x + mm * _lin(hdn, sd[p + 'mlp.fc2.weight'], sd[p + 'mlp.fc2.bias'])


def cls_aggregator(sd, prefix: str, x: Tensor) -> Tensor:
    """BaseEncoderLayer.forward (motionformer.py:301-334) around nn.TransformerEncoderLayer(norm_first=True,
    nhead 12, ff 3072, exact GELU, eps 1e-6, dropout 0): prepend learned CLS, one pre-norm layer, return row 0.
    x (G, L, 768) -> (G, 768).  Computed densely (all rows), as the reference does."""
    G, L, _ = x.shape
    h, d = 12, D // 12
    x = torch.cat([sd[prefix + 'cls_token'].expand(G, 1, D), x], dim=1)
    y = _ln(x, sd[prefix + 'norm1.weight'], sd[prefix + 'norm1.bias'], EPS_V)
    qkv = _lin(y, sd[prefix + 'self_attn.in_proj_weight'], sd[prefix + 'self_attn.in_proj_bias'])
    q, k, v = [t.reshape(G, L + 1, h, d).permute(0, 2, 1, 3) for t in qkv.chunk(3, dim=-1)]
    a = _softmax_av(q * d ** -0.5, k, v).permute(0, 2, 1, 3).reshape(G, L + 1, D)
    x = x + _lin(a, sd[prefix + 'self_attn.out_proj.weight'], sd[prefix + 'self_attn.out_proj.bias'])
    y = _ln(x, sd[prefix + 'norm2.weight'], sd[prefix + 'norm2.bias'], EPS_V)
    y = _lin(_gelu(_lin(y, sd[prefix + 'linear1.weight'], sd[prefix + 'linear1.bias'])),
             sd[prefix + 'linear2.weight'], sd[prefix + 'linear2.bias'])
    return (x + y)[:, 0]


def extract_vfeats(sd, vis: Tensor, dtype=torch.float32, taps: Optional[dict] = None, device='cpu') -> Tensor:
    """Synchformer.extract_vfeats (sync_model.py:72-80) -> MotionFormer.forward/forward_segments
    (motionformer.py:182-252).  vis (B, S, 16, 3, 224, 224) -> (B, S, 8, 768)."""
    sd = _cast_sd({k: v for k, v in sd.items() if k.startswith('vfeat_extractor.')}, dtype, device)
    B, S = vis.shape[:2]
    x = video_patch_embed(sd, vis.reshape(B * S, *vis.shape[2:]).to(device, dtype))
    if taps is not None:
        taps['v_embed'] = x
    for i in range(V_DEPTH):
        x = video_block(sd, i, x)
        if taps is not None and i in (0, V_DEPTH - 1):
            taps[f'v_block{i}'] = x
    x = _ln(x[:, 1:], sd['vfeat_extractor.norm.weight'], sd['vfeat_extractor.norm.bias'], EPS_V)   # :229-232
    x = x.reshape(B * S * V_FRAMES, V_SPACE, D)                    # :254-272 + :363  '(BS t) (h w) D'
    x = cls_aggregator(sd, 'vfeat_extractor.spatial_attn_agg.', x)
    return x.reshape(B, S, V_FRAMES, D)


# ----------------------------------------------------------------------------------------------------------
# a9-a13: audio stream
# ----------------------------------------------------------------------------------------------------------
def audio_patch_embed(sd, spec: Tensor) -> Tensor:
    """spec (BS, 128, 66) [freq, time] -> (BS, 74, 768).  modeling_ast.py:113-117 (Conv2d k16, stride 10 over
    (freq, time)) + :83-93 (cls, distillation, + position embeddings).  Token order 2 + f*6 + t."""
    p = 'afeat_extractor.ast.embeddings.'
    BS = spec.shape[0]
    patches = spec.unfold(1, 16, 10).unfold(2, 16, 10)                            # (BS, 12, 6, 16, 16)
    tok = _lin(patches.reshape(BS, A_F * A_T, 256),
               sd[p + 'patch_embeddings.projection.weight'].reshape(D, 256), sd[p + 'patch_embeddings.projection.bias'])
    x = torch.cat([sd[p + 'cls_token'].expand(BS, 1, D), sd[p + 'distillation_token'].expand(BS, 1, D), tok], dim=1)
    return x + sd[p + 'position_embeddings']


def ast_layer(sd, i: int, x: Tensor) -> Tensor:
    """ASTLayer.forward modeling_ast.py:294-322 with ASTSelfAttention :145-184 (scores / sqrt(64))."""
    p = f'afeat_extractor.ast.encoder.layer.{i}.'
    BS, N, _ = x.shape
    h, d = A_HEADS, D // A_HEADS
    y = _ln(x, sd[p + 'layernorm_before.weight'], sd[p + 'layernorm_before.bias'], EPS_A)
    sp = lambda t: t.reshape(BS, N, h, d).permute(0, 2, 1, 3)
    q = sp(_lin(y, sd[p + 'attention.attention.query.weight'], sd[p + 'attention.attention.query.bias']))
    k = sp(_lin(y, sd[p + 'attention.attention.key.weight'], sd[p + 'attention.attention.key.bias']))
    v = sp(_lin(y, sd[p + 'attention.attention.value.weight'], sd[p + 'attention.attention.value.bias']))
    a = _softmax_av(q / math.sqrt(d), k, v).permute(0, 2, 1, 3).reshape(BS, N, D)
    x = x + _lin(a, sd[p + 'attention.output.dense.weight'], sd[p + 'attention.output.dense.bias'])
    y = _ln(x, sd[p + 'layernorm_after.weight'], sd[p + 'layernorm_after.bias'], EPS_A)
    y = _gelu(_lin(y, sd[p + 'intermediate.dense.weight'], sd[p + 'intermediate.dense.bias']))
    return x + _lin(y, sd[p + 'output.dense.weight'], sd[p + 'output.dense.bias'])


def extract_afeats(sd, aud: Tensor, dtype=torch.float32, taps: Optional[dict] = None, device='cpu') -> Tensor:
    """Synchformer.extract_afeats (sync_model.py:82-89) -> AST.forward/forward_segments (ast.py:137-201).
    aud (B, S, 1, 128, 66) -> (B, S, 6, 768).  The two transposes (sync_model.py:84, modeling_ast.py:115) cancel."""
    sd = _cast_sd({k: v for k, v in sd.items() if k.startswith('afeat_extractor.')}, dtype, device)
    B, S = aud.shape[:2]
    x = audio_patch_embed(sd, aud.reshape(B * S, 128, 66).to(device, dtype))
    if taps is not None:
        taps['a_embed'] = x
    for i in range(A_DEPTH):
        x = ast_layer(sd, i, x)
    x = _ln(x, sd['afeat_extractor.ast.layernorm.weight'], sd['afeat_extractor.ast.layernorm.bias'], EPS_A)  # :543
    if taps is not None:
        taps['a_last_hidden'] = x
    x = x[:, 2:].reshape(B * S, A_F, A_T, D).permute(0, 2, 1, 3).reshape(B * S * A_T, A_F, D)   # ast.py:215-238, 266-268
    x = cls_aggregator(sd, 'afeat_extractor.freq_attn_agg.', x)
    return x.reshape(B, S, A_T, D)


# ----------------------------------------------------------------------------------------------------------
# a14-a16: projections + synchronisation transformer + loss
# ----------------------------------------------------------------------------------------------------------
def sync_block(sd, i: int, x: Tensor) -> Tensor:
    """Block.forward modules/transformer.py:94-97 with SelfAttention.forward :58-76 (8 heads x 96, scores * 96^-0.5)."""
    p = f'transformer.blocks.{i}.'
    B, T, _ = x.shape
    h, d = S_HEADS, D // S_HEADS
    y = _ln(x, sd[p + 'ln1.weight'], sd[p + 'ln1.bias'], EPS_S)
    sp = lambda t: t.reshape(B, T, h, d).permute(0, 2, 1, 3)
    q = sp(_lin(y, sd[p + 'attn.query.weight'], sd[p + 'attn.query.bias']))
    k = sp(_lin(y, sd[p + 'attn.key.weight'], sd[p + 'attn.key.bias']))
    v = sp(_lin(y, sd[p + 'attn.value.weight'], sd[p + 'attn.value.bias']))
    a = _softmax_av(q * (1.0 / math.sqrt(d)), k, v).permute(0, 2, 1, 3).reshape(B, T, D)
    x = x + _lin(a, sd[p + 'attn.proj.weight'], sd[p + 'attn.proj.bias'])
    y = _ln(x, sd[p + 'ln2.weight'], sd[p + 'ln2.bias'], EPS_S)
    y = _gelu(_lin(y, sd[p + 'mlp.0.weight'], sd[p + 'mlp.0.bias']))
    return x + _lin(y, sd[p + 'mlp.2.weight'], sd[p + 'mlp.2.bias'])


def sync_head(sd, vfeat: Tensor, afeat: Tensor, dtype=torch.float32, head: str = 'off_head', device='cpu') -> Tensor:
    """sync_model.py:55-62 (vproj/aproj, flatten segments) + GlobalTransformer.forward :150-173.
    vfeat (B,S,8,768), afeat (B,S,6,768) -> logits (B, 21).  head='sync_head' gives the 2-class variant
    (GlobalTransformerWithSyncabilityHead :176-190)."""
    sd = _cast_sd({k: v for k, v in sd.items() if k.split('.')[0] in ('vproj', 'aproj', 'transformer')}, dtype, device)
    B, S = vfeat.shape[:2]
    v = _lin(vfeat.to(device, dtype), sd['vproj.weight'], sd['vproj.bias']).reshape(B, S * 8, D)
    a = _lin(afeat.to(device, dtype), sd['aproj.weight'], sd['aproj.bias']).reshape(B, S * 6, D)
    t = 'transformer.'
    v = _ln(v, sd[t + 'vis_in_lnorm.weight'], sd[t + 'vis_in_lnorm.bias'], EPS_S)
    a = _ln(a, sd[t + 'aud_in_lnorm.weight'], sd[t + 'aud_in_lnorm.bias'], EPS_S)
    x = torch.cat([sd[t + 'OFF_tok'].expand(B, 1, D), v, sd[t + 'MOD_tok'].expand(B, 1, D), a], dim=1)
    pos = sd[t + 'pos_emb_cfg.pos_emb']
    assert pos.shape[1] == x.shape[1], f'pos_emb length {pos.shape[1]} != sequence {x.shape[1]} (transformer.py:129-130 adds the full table)'
    x = x + pos
    for i in range(S_DEPTH):
        x = sync_block(sd, i, x)
    x = _ln(x, sd[t + 'ln_f.weight'], sd[t + 'ln_f.bias'], EPS_S)
    return _lin(x[:, 0], sd[t + head + '.weight'], sd[t + head + '.bias'])


def forward(sd, vis: Tensor, aud: Tensor, targets: Optional[Tensor] = None, dtype=torch.float32, taps: Optional[dict] = None, device='cpu'):
    """Synchformer.forward sync_model.py:38-70 -> (loss | None, logits (B, 21)).  `device` exists so that tests can run the same
    restatement with torch's CUDA library kernels as a second opinion; the oracle proper is the CPU run."""
    vf = extract_vfeats(sd, vis, dtype, taps, device)
    af = extract_afeats(sd, aud, dtype, taps, device)
    if taps is not None:
        taps['vfeats'], taps['afeats'] = vf, af
    logits = sync_head(sd, vf, af, dtype, device=device)
    loss = None
    if targets is not None:                                   # compute_loss sync_model.py:91-99
        loss = F.cross_entropy(logits.float(), targets.to(device))
    return loss, logits


def avclip_features(sd, vis: Tensor, aud: Tensor, dtype=torch.float32):
    """Stage-I encode_stream (open_clip/model.py:536-545) with agg_time_module='AveragePooling'
    (motionformer.py:395-409): mean over the 8 / 6 time tokens, identity bridge, L2-normalise -> (BS,768) x2."""
    vf = extract_vfeats(sd, vis, dtype).mean(2).reshape(-1, D)
    af = extract_afeats(sd, aud, dtype).mean(2).reshape(-1, D)
    return F.normalize(vf, dim=-1), F.normalize(af, dim=-1)


def avclip_loss(vfeat: Tensor, afeat: Tensor, scale, vfeat_all: Optional[Tensor] = None, afeat_all: Optional[Tensor] = None) -> Tensor:
    """AVCLIP.compute_loss (open_clip/model.py:507-527) on L2-normalised (n, D) features: similarities of the local rows against the
    (optionally all-gathered, :492-494) rows, divided by the temperature `scale`, soft targets `eye(n, N)` exactly as `_make_targets`
    builds them (:515-522: the positive of local row i is column i whatever the rank), mean of the two cross-entropies."""
    vfeat_all = vfeat if vfeat_all is None else vfeat_all
    afeat_all = afeat if afeat_all is None else afeat_all
    sim_v2a = vfeat @ afeat_all.mT / scale
    sim_a2v = afeat @ vfeat_all.mT / scale
    tgt = torch.eye(*sim_v2a.shape, dtype=sim_v2a.dtype)
    return (F.cross_entropy(sim_v2a, tgt) + F.cross_entropy(sim_a2v, tgt)) / 2


# ----------------------------------------------------------------------------------------------------------
# N3: training-mode forward of the synchronisation module with EXPLICIT dropout multipliers, so that torch
# autograd on this restatement gives the reference gradients for a known mask (scripts/train_utils.py:373-386 drives
# loss.backward() through model/sync_model.py:55-62, 150-173 and modules/transformer.py:58-97).
# Pinned by tests/golden/sync_train_b2s2.npz (the reference's own modules in train mode with the same multipliers injected in
# place of its nn.Dropout instances; tests/golden/make_golden_train.py).
# ----------------------------------------------------------------------------------------------------------
def train_sites(n_layer: int = S_DEPTH) -> dict:
    """Dropout layer -> site id of the counter-based mask (csrc/philox.cuh)."""
    sites = {'embd': 0}
    for i in range(n_layer):
        sites[f'attn{i}'], sites[f'resid_attn{i}'], sites[f'resid_mlp{i}'] = 1 + 3 * i, 2 + 3 * i, 3 + 3 * i
    return sites


def train_multipliers(B: int, T: int, seed: int, embd_pdrop: float = 0.1, resid_pdrop: float = 0.1, attn_pdrop: float = 0.1) -> Dict[str, Tensor]:
    """The multipliers (0 or 1 / (1 - p)) the CUDA kernels apply for `seed`: embd / resid_* (B, T, 768), attn* (B, 8, T, T)."""
    from . import philox
    out = {}
    for name, site in train_sites().items():
        if name == 'embd':
            shape, p = (B, T, D), embd_pdrop
        elif name.startswith('attn'):
            shape, p = (B, S_HEADS, T, T), attn_pdrop
        else:
            shape, p = (B, T, D), resid_pdrop
        out[name] = torch.from_numpy(philox.dropout_multiplier(shape, p, seed, site))
    return out


def sync_block_train(sd, i: int, x: Tensor, mult: Optional[Dict[str, Tensor]]) -> Tensor:
    """Block.forward transformer.py:94-97 in training mode: attn_drop on the softmax output (:74), resid_drop after proj (:76)
    and after the MLP (:92).  mult=None -> dropout off."""
    p = f'transformer.blocks.{i}.'
    B, T, _ = x.shape
    h, d = S_HEADS, D // S_HEADS
    m = (lambda k: 1.0) if mult is None else (lambda k: mult[k].to(x.dtype))
    y = _ln(x, sd[p + 'ln1.weight'], sd[p + 'ln1.bias'], EPS_S)
    sp = lambda t: t.reshape(B, T, h, d).permute(0, 2, 1, 3)
    q = sp(_lin(y, sd[p + 'attn.query.weight'], sd[p + 'attn.query.bias']))
    k = sp(_lin(y, sd[p + 'attn.key.weight'], sd[p + 'attn.key.bias']))
    v = sp(_lin(y, sd[p + 'attn.value.weight'], sd[p + 'attn.value.bias']))
    att = torch.softmax((q @ k.transpose(-1, -2)) * (1.0 / math.sqrt(d)), dim=-1) * m(f'attn{i}')
    a = (att @ v).permute(0, 2, 1, 3).reshape(B, T, D)
    x = x + _lin(a, sd[p + 'attn.proj.weight'], sd[p + 'attn.proj.bias']) * m(f'resid_attn{i}')
    y = _ln(x, sd[p + 'ln2.weight'], sd[p + 'ln2.bias'], EPS_S)
    y = _gelu(_lin(y, sd[p + 'mlp.0.weight'], sd[p + 'mlp.0.bias']))
    return x + _lin(y, sd[p + 'mlp.2.weight'], sd[p + 'mlp.2.bias']) * m(f'resid_mlp{i}')


def sync_head_train(sd, vfeat: Tensor, afeat: Tensor, mult: Optional[Dict[str, Tensor]] = None, head: str = 'off_head') -> Tensor:
    """sync_head() in training mode; `sd` tensors may require grad (no dtype cast is applied: pass fp32 / fp64 leaves)."""
    B, S = vfeat.shape[:2]
    v = _lin(vfeat, sd['vproj.weight'], sd['vproj.bias']).reshape(B, S * 8, D)
    a = _lin(afeat, sd['aproj.weight'], sd['aproj.bias']).reshape(B, S * 6, D)
    t = 'transformer.'
    v = _ln(v, sd[t + 'vis_in_lnorm.weight'], sd[t + 'vis_in_lnorm.bias'], EPS_S)
    a = _ln(a, sd[t + 'aud_in_lnorm.weight'], sd[t + 'aud_in_lnorm.bias'], EPS_S)
    x = torch.cat([sd[t + 'OFF_tok'].expand(B, 1, D), v, sd[t + 'MOD_tok'].expand(B, 1, D), a], dim=1)
    x = x + sd[t + 'pos_emb_cfg.pos_emb']
    if mult is not None:
        x = x * mult['embd'].to(x.dtype)                       # self.drop sync_model.py:168
    for i in range(S_DEPTH):
        x = sync_block_train(sd, i, x, mult)
    x = _ln(x, sd[t + 'ln_f.weight'], sd[t + 'ln_f.bias'], EPS_S)
    return _lin(x[:, 0], sd[t + head + '.weight'], sd[t + head + '.bias'])


def sync_train_grads(sd, vfeat: Tensor, afeat: Tensor, targets: Tensor, mult: Optional[Dict[str, Tensor]] = None, head: str = 'off_head',
                     dtype=torch.float32, loss_scale: float = 1.0):
    """loss = cross_entropy(logits, targets) (compute_loss sync_model.py:91-99); returns (loss, logits, {name: d loss / d param}) for
    every parameter of vproj / aproj / transformer (the trainable set of configs/sync.yaml with frozen extractors)."""
    names = [k for k in sd if k.split('.')[0] in ('vproj', 'aproj', 'transformer')]
    leaves = {k: sd[k].detach().to(dtype).clone().requires_grad_(True) for k in names}
    with torch.enable_grad():
        logits = sync_head_train(leaves, vfeat.to(dtype), afeat.to(dtype), mult, head)
        loss = F.cross_entropy(logits, targets)
        grads = torch.autograd.grad(loss * loss_scale, [leaves[k] for k in names])
    return loss.detach(), logits.detach(), {k: g for k, g in zip(names, grads)}


# ----------------------------------------------------------------------------------------------------------
# N1: stage-I training step of the two encoders - AVCLIP.forward / compute_loss
# (train_clip_src/open_clip/model.py:474-527) with both towers in train mode.  The only stochastic layer is the Motionformer's
# DropPath (rate 0.2 i / 11 in block i, divided_224_16x4.yaml:59, video_model_builder.py:86-87); its per-segment multipliers are
# explicit inputs here.  Pinned by tests/golden/encoders_train_b1s2.npz (the reference towers with the same multipliers injected in
# place of their DropPath instances; tests/golden/make_golden_encoders_train.py).
# ----------------------------------------------------------------------------------------------------------
DROP_PATH_RATE = 0.2


def drop_path_multipliers(n_segments: int, seed: int, rate: float = DROP_PATH_RATE) -> dict:
    """{block i: (space (n, 1, 1), mlp (n, 1, 1))}: what the CUDA kernels apply for `seed` (sites 2 i and 2 i + 1, one decision per
    segment, csrc/philox.cuh)."""
    from . import philox
    out = {}
    for i in range(V_DEPTH):
        p = rate * i / (V_DEPTH - 1)
        out[i] = tuple(torch.from_numpy(philox.dropout_multiplier((n_segments,), p, seed, 2 * i + k)).reshape(n_segments, 1, 1) for k in (0, 1))
    return out


def encoder_features_train(sd, vis: Tensor, aud: Tensor, drop_path: Optional[dict] = None):
    """Differentiable towers on the given leaves (no cast, no detach): vis (B, S, 16, 3, 224, 224), aud (B, S, 1, 128, 66) ->
    ((B, S, 8, 768), (B, S, 6, 768))."""
    B, S = vis.shape[:2]
    x = video_patch_embed(sd, vis.reshape(B * S, *vis.shape[2:]))
    for i in range(V_DEPTH):
        x = video_block(sd, i, x, None if drop_path is None else drop_path[i])
    x = _ln(x[:, 1:], sd['vfeat_extractor.norm.weight'], sd['vfeat_extractor.norm.bias'], EPS_V)
    v = cls_aggregator(sd, 'vfeat_extractor.spatial_attn_agg.', x.reshape(B * S * V_FRAMES, V_SPACE, D)).reshape(B, S, V_FRAMES, D)
    y = audio_patch_embed(sd, aud.reshape(B * S, 128, 66))
    for i in range(A_DEPTH):
        y = ast_layer(sd, i, y)
    y = _ln(y, sd['afeat_extractor.ast.layernorm.weight'], sd['afeat_extractor.ast.layernorm.bias'], EPS_A)
    y = y[:, 2:].reshape(B * S, A_F, A_T, D).permute(0, 2, 1, 3).reshape(B * S * A_T, A_F, D)
    a = cls_aggregator(sd, 'afeat_extractor.freq_attn_agg.', y).reshape(B, S, A_T, D)
    return v, a


def contrastive_loss(vfeat: Tensor, afeat: Tensor, logit_scale) -> Tensor:
    """AVCLIP.compute_loss / _loss (open_clip/model.py:507-527) on time-pooled (B, S, 8|6, 768) features: mean over the time tokens
    (AveragePooling, motionformer.py:395-409), flatten (B S), L2-normalise, symmetric soft-target cross-entropy with identity targets."""
    v = F.normalize(vfeat.mean(2).reshape(-1, D), dim=-1)
    a = F.normalize(afeat.mean(2).reshape(-1, D), dim=-1)
    sim_v2a, sim_a2v = v @ a.mT / logit_scale, a @ v.mT / logit_scale
    tgt = torch.eye(*sim_v2a.shape, dtype=sim_v2a.dtype)
    return (F.cross_entropy(sim_v2a, tgt) + F.cross_entropy(sim_a2v, tgt)) / 2


def encoders_train_grads(sd, vis: Tensor, aud: Tensor, drop_path: Optional[dict] = None, logit_scale: float = 0.07, dtype=torch.float32):
    """-> (loss, vfeat, afeat, {name: d loss / d param}) for every extractor parameter that takes part in the forward."""
    names = [k for k in sd if k.split('.')[0] in ('vfeat_extractor', 'afeat_extractor') and '.patch_embed.proj.' not in k]
    leaves = {k: sd[k].detach().to(dtype).clone().requires_grad_(True) for k in names}
    with torch.enable_grad():
        v, a = encoder_features_train(leaves, vis.to(dtype), aud.to(dtype), drop_path)
        loss = contrastive_loss(v, a, logit_scale)
        grads = torch.autograd.grad(loss, [leaves[k] for k in names])
    return loss.detach(), v.detach(), a.detach(), dict(zip(names, grads))


def shift_and_get_preds(a: Tensor, v: Tensor, W: int):
    """training/train.py:549-579 restated: sliding windows of W segments over (B, S, D) features, all-pairs window similarity, top-1 in
    both directions -> (preds_a, preds_v) each (B, S - W + 1); also returns the similarity matrix for tolerance-aware comparisons."""
    B, S, D_ = a.shape
    af = a.unfold(-2, W, 1).contiguous().view(B, S - W + 1, D_ * W)
    vf = v.unfold(-2, W, 1).contiguous().view(B, S - W + 1, D_ * W)
    sim = af @ vf.mT
    return torch.argmax(sim, dim=-2), torch.argmax(sim, dim=-1), sim
