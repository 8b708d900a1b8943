/*
 * synchformer_b200 — C-ABI of the B200 (sm_100a) kernels behind `model.sync_model.Synchformer.forward`.
 *
 * The reference (v-iashin/Synchformer) is pure Python/PyTorch and has no FFI of its own: every entry point
 * below replaces a group of torch ops on the hot path and cites the reference lines it stands in for
 * (paths relative to the reference root).  The host side (`synchformer_b200/model.py`) mirrors the
 * reference's `Synchformer` constructor / forward / state_dict surface and calls these through ctypes.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless stated otherwise
 *   - `bf16` buffers are passed as `void*` (raw __nv_bfloat16), fp32 as `float*`
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), never allocates device
 *     memory, never synchronises; the caller owns all buffers
 *   - return value: 0 = ok, negative = SFB_E_* (no exceptions cross the ABI); `sfb_last_error()` gives text
 *   - leading dimensions / strides are in ELEMENTS
 */
#ifndef SYNCHFORMER_B200_H
#define SYNCHFORMER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SFB_OK 0
#define SFB_E_INVALID (-1)   /* bad argument (shape / alignment / null pointer) */
#define SFB_E_CUDA (-2)      /* a CUDA runtime / driver call failed */
#define SFB_E_UNSUPPORTED (-3)

/* library / device probing (host-only, no kernels) */
int sfb_abi_version(void);
const char *sfb_last_error(void);
/* 0 if the current device can run the kernels (compute capability 10.x), SFB_E_UNSUPPORTED otherwise */
int sfb_device_check(void);

/* ---------------------------------------------------------------------------------------------------------
 * K5 — every nn.Linear on the path (vit_helper.py:105,156,393-396; modeling_ast.py:149-152,200,258,272;
 * modules/transformer.py:62-64,74,87-92; motionformer.py:329 via nn.TransformerEncoderLayer; sync_model.py:55-56)
 *
 *   out[M,N] = epilogue( A[M,K] (bf16, row stride lda) x W[N,K]^T (bf16, row stride K) + bias[N] (fp32) )
 *
 * epilogue flags: SFB_GEMM_GELU      exact erf GELU (nn.GELU / HF GELUActivation)
 *                 SFB_GEMM_RESIDUAL  += residual[M,N] fp32 (row stride ldr; ldr == 0 broadcasts one row)
 *                 SFB_GEMM_OUT_F32   write fp32 instead of bf16 (out may alias residual)
 * tcgen05.mma (fp32 accumulators in TMEM), TMA-fed shared-memory ring, persistent CTAs; CTA pairs (cta_group::2,
 * 256x256 tiles) for M > 128, single CTAs (128x256 tiles) otherwise.  RESIDUAL requires OUT_F32.
 * Requirements: K % 8 == 0, lda % 8 == 0, N % 8 == 0, ldo % 8 == 0, A/W 16-byte aligned.
 * `impl`: 0 = product path (auto); 1 = plain CUDA-core kernel kept only as a bring-up cross-check;
 *         2 / 3 = force the single-CTA / CTA-pair tcgen05 kernel (tests, A/B measurements). */
#define SFB_GEMM_GELU 1
#define SFB_GEMM_RESIDUAL 2
#define SFB_GEMM_OUT_F32 4
#define SFB_GEMM_EMIT_LN 8
#define SFB_GEMM_LN_FOLD 16
int sfb_gemm_bf16(const void *A, int64_t lda, const void *W, const float *bias, const float *residual, int64_t ldr,
                  void *out, int64_t ldo, int M, int N, int K, int flags, int impl, void *stream);

/* K5 + K4 fused: nn.LayerNorm folded into the Linear on either side of it (DividedSpaceTimeBlock.forward vit_helper.py:364-376:
 * norm3 -> timeattn.qkv, norm1 -> attn.qkv, norm2 -> mlp.fc1, each fed by the residual update before it).  Same as sfb_gemm_bf16 plus
 *   SFB_GEMM_EMIT_LN  (with RESIDUAL | OUT_F32, N % 64 == 0): also writes emit_bf16[M, N] = bf16(out) (row stride ld_emit) and
 *                     emit_stats[M][N / 64][2] = (sum, sum of squares) of the fp32 out row over each 64-column group
 *   SFB_GEMM_LN_FOLD  (bf16 output): A holds UN-normalised rows, W = bf16(gamma . W0), bias = b0 + W0 beta, ln_colsum[N] = row sums of W
 *                     as stored (bf16 values, fp32 sum), ln_stats[M][ln_parts][2] partial (sum, sum of squares) of the A rows over K;
 *                         out = epilogue( rstd (A W^T - mean ln_colsum) + bias ),  mean = sum / K,  rstd = rsqrt(sumsq / K - mean^2 + ln_eps)
 *                     which equals LayerNorm(A; gamma, beta, ln_eps) W0^T + b0 with bf16 operand rounding applied to A instead of to
 *                     the normalised rows. */
int sfb_gemm_bf16_ln(const void *A, int64_t lda, const void *W, const float *bias, const float *residual, int64_t ldr,
                     void *out, int64_t ldo, int M, int N, int K, int flags, int impl, const float *ln_stats, int ln_parts,
                     const float *ln_colsum, float ln_eps, float *emit_stats, void *emit_bf16, int64_t ld_emit, void *stream);
/* x (rows, 768) fp32 (row stride ldx) -> xb (rows, 768) bf16 contiguous and stats[rows][1][2] = (sum, sum of squares) per row: the
 * LN_FOLD inputs for a residual stream that no EMIT_LN GEMM produced (the token assembly before block 0). */
int sfb_rowstats_cast(const float *x, int64_t ldx, void *xb, float *stats, int rows, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * K4 — nn.LayerNorm over D = 768 with fp32 statistics (eps 1e-6 Motionformer/aggregators
 * video_model_builder.py:39, motionformer.py:126; 1e-12 AST modeling_ast.py:291-292,464; 1e-5 sync
 * transformer.py:84-85, sync_model.py:126-127,143).
 * Output row r reads input row  (r / group) * group_stride + offset + (r % group)   (drops CLS / aux tokens
 * without a copy: motionformer.py:229-232, ast.py:232-233).  If gamma2 != NULL a second LayerNorm
 * (gamma2, beta2, eps2) is applied to the result of the first in registers (final encoder norm followed by the
 * aggregator's norm1, motionformer.py:231 + nn.TransformerEncoderLayer norm_first). */
int sfb_layernorm(const float *x, int64_t ldx, void *out, int64_t ldo, int out_f32, const float *gamma, const float *beta,
                  float eps, const float *gamma2, const float *beta2, float eps2, int rows, int group, int group_stride,
                  int offset, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * K6-K11 — multi-head softmax attention on strided bf16 Q/K/V views (no rearrange/cat copies).
 * A "problem" is (outer o, inner i, head h); its rows live at
 *     q   + o*q_outer  + i*q_inner  + h*head_dim + r*q_row          r in [0, Lq)
 *     k/v + o*kv_outer + i*kv_inner + h*head_dim + r*kv_row         r in [0, Lk)
 *     out + o*o_outer  + i*o_inner  + h*head_dim + r*o_row
 * and, if k_prefix != NULL, one extra key/value row (the CLS token, vit_helper.py:129-134; the aggregator CLS,
 * motionformer.py:305-306) at  k_prefix/v_prefix + o*prefix_outer + h*head_dim  is placed BEFORE the Lk rows.
 * softmax(scale * q k^T) v with fp32 scores/statistics.
 * Covers: time attention (8 x 9, vit_helper.py:100-158 '(b n) f d'), space attention (196 x 197, '(b f) n d'),
 * Motionformer CLS query (1 x 1569, vit_helper.py:124), AST MHSA (74 x 74, modeling_ast.py:145-184),
 * CLS-aggregator attention rows (1 x 197 / 1 x 13, motionformer.py:329), sync MHSA (198 x 198, hd 96,
 * modules/transformer.py:58-76).  head_dim in {64, 96}. */
typedef struct sfb_attn_desc {
    const void *q, *k, *v;
    const void *k_prefix, *v_prefix; /* NULL = no prefix row */
    void *out;
    int64_t q_outer, q_inner, q_row;
    int64_t kv_outer, kv_inner, kv_row;
    int64_t o_outer, o_inner, o_row;
    int64_t prefix_outer;
    int32_t n_outer, n_inner, n_heads, head_dim, Lq, Lk;
    float scale;
    int32_t impl; /* 0 = auto (specialised kernels), 1 = generic CUDA-core kernel (bring-up cross-check) */
    /* Optional fused extra query (the Motionformer CLS query, vit_helper.py:124, which attends to ALL keys of its segment):
     * one more query row per (outer, head) at  q_extra + o*q_extra_outer + h*head_dim  rides along with every inner problem of
     * that outer index; its softmax state over that problem's keys (the prefix key is counted for inner == 0 only) goes to
     *   extra_partial[((o*n_heads + h)*n_inner + i)*(head_dim + 2)] = { max (log2 units), sum, out[head_dim] / sum }   (fp32)
     * and sfb_attention_merge_partials() combines the n_inner states.  NULL = off.  Supported where
     * sfb_attention_extra_supported() says so (the tcgen05 space-attention kernel; with SFB_TIME_CLS_FUSED=1 also the time-attention kernel). */
    const void *q_extra;
    int64_t q_extra_outer;
    float *extra_partial;
} sfb_attn_desc;
int sfb_attention(const sfb_attn_desc *desc, void *stream);
/* 1 if sfb_attention(desc) would honour q_extra / extra_partial for this descriptor, 0 otherwise (host-only, no launch) */
int sfb_attention_extra_supported(const sfb_attn_desc *desc);
/* out[o*out_outer + h*head_dim + d] (bf16) = sum_i w_i partial_out_i[d] / sum_i w_i,  w_i = sum_i * 2^(max_i - max) */
int sfb_attention_merge_partials(const float *partial, void *out, int64_t out_outer, int n_outer, int n_inner, int n_heads, int head_dim,
                                 void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * K2 — PatchEmbed3D as im2col + GEMM (vit_helper.py:436-445).  vis is (n_seg, 16, 3, 224, 224) in the layout
 * Synchformer.forward receives it (sync_model.py:38-46; the permute at :74 is only a view).
 * in_dtype: 0 fp32, 1 fp16, 2 bf16, 3 uint8 (uint8 fuses RGBToHalfToZeroOne + RGBNormalize(.5,.5),
 * dataset/transforms.py:647-669).  A is (n_seg*1568, 1536) bf16, K order (c, dt, dy, dx) = Conv3d weight order. */
int sfb_im2col_video(const void *vis, int in_dtype, void *A, int n_seg, void *stream);
/* N2: segment slicing fused into the gather - GenerateMultipleSegments (dataset/transforms.py:402-499) on the device.
 * clip is (n_clips, n_frames, 3, 224, 224); segment s of clip b covers frames [v_start + s*v_stride, +16); A is
 * (n_clips*n_segments*1568, 1536).  Overlapping segments (stride 8) are no longer shipped twice over PCIe. */
int sfb_im2col_video_clip(const void *clip, int in_dtype, void *A, int n_clips, int n_frames, int n_segments, int v_start, int v_stride,
                          void *stream);
/* + bias already added by the GEMM; adds pos_embed[1+n] + temp_embed[f], prepends cls_token + pos_embed[0]
 * (video_model_builder.py:221-254).  patch (n_seg*1568, 768) fp32 -> x (n_seg, 1569, 768) fp32 */
int sfb_video_tokens(const float *patch, const float *cls_token, const float *pos_embed, const float *temp_embed, float *x,
                     int n_seg, void *stream);

/* K3 — ASTPatchEmbeddings + ASTEmbeddings (modeling_ast.py:113-117, 83-93).  spec (n_seg, 128, 66) fp32
 * [freq, time]; A (n_seg*72, 256) bf16; tokens 2 + f*6 + t. */
int sfb_im2col_ast(const float *spec, void *A, int n_seg, void *stream);
int sfb_ast_tokens(const float *patch, const float *cls_token, const float *dist_token, const float *pos_embed, float *x,
                   int n_seg, void *stream);

/* K12 — GlobalTransformer token assembly (sync_model.py:150-167, modules/transformer.py:129-130):
 * x[b] = [OFF, LN_vis(v[b, 0..8S)), MOD, LN_aud(a[b, 0..6S))] + pos_emb;  v (B, 8S, 768), a (B, 6S, 768) fp32 */
int sfb_sync_tokens(const float *v, const float *a, const float *vis_ln_w, const float *vis_ln_b, const float *aud_ln_w,
                    const float *aud_ln_b, float eps, const float *off_tok, const float *mod_tok, const float *pos_emb, float *x,
                    int B, int S, void *stream);
/* K13 — ln_f on token 0 + Linear(768 -> n_cls) in fp32 (sync_model.py:169-172).  x (B, T, 768) fp32 */
int sfb_sync_head(const float *x, int T, const float *ln_w, const float *ln_b, float eps, const float *W, const float *b,
                  float *logits, int B, int n_cls, void *stream);

/* fp32 -> bf16 cast of n elements (n % 4 == 0); used for weights at load time and the aggregator features */
int sfb_cast_f32_bf16(const float *in, void *out, int64_t n, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * K1 — mel front-end (dataset/transforms.py:815-871 = torchaudio MelSpectrogram(n_fft 1024, win 400, hop 160,
 * 128 mels, power 2) -> log(x + 1e-6) -> pad T 65->66 with 0.0 -> (x + 4.2677393) / (2 * 4.5689974)).
 * wave (n_seg, 10240) fp32 -> out (n_seg, 128, 66) fp32.  The first call uploads three constant tables
 * (twiddles, window, filterbank) to static device arrays with a synchronous copy. */
int sfb_mel_frontend(const float *wave, float *out, int n_seg, void *stream);
/* same on un-duplicated waveforms: wave (n_clips, clip_stride samples); segment s of clip b = samples [a_start + s*a_stride, +10240) */
int sfb_mel_frontend_clip(const float *wave, int64_t clip_stride, float *out, int n_clips, int n_segments, int a_start, int a_stride,
                          void *stream);

/* =========================================================================================================
 * N3 — training step of the synchronisation module: vproj / aproj + GlobalTransformer forward with dropout and
 * backward (modules/transformer.py:31-97, sync_model.py:55-62, 150-173; driven by scripts/train_utils.py:373-386).  The linear
 * layers' dX = dY W and dW = dY^T X reuse sfb_gemm_bf16 on operands transposed by sfb_transpose_bf16; everything else is below.
 * ========================================================================================================= */

/* nn.Dropout (transformer.py:47-48,74,92; sync_model.py:137) with a counter-based mask, forward and backward in one entry point:
 *     out[e] = (residual ? residual[e] : 0) + in[e] * (keep(seed, site, e) ? 1 / (1 - p) : 0)         e in [0, n), n % 4 == 0
 *     keep(seed, site, e) = Philox4x32-10(counter = (e >> 2, site, 0), key = seed)[e & 3] >= floor(p * 2^32)
 * in / residual fp32, out fp32 (out_bf16 == 0) or bf16; out may alias in or residual.  p == 0 degenerates to a copy / cast / add.
 * The same (seed, site, e) gives the same mask in every kernel of the library; e is the linear index into the dropped tensor. */
int sfb_dropout(const float *in, const float *residual, void *out, int out_bf16, int64_t n, float p, uint64_t seed, uint32_t site,
                void *stream);

/* nn.GELU on the bf16 pre-activation (transformer.py:89) and its derivative: y = gelu(x); dx = dy * (Phi(x) + x phi(x)).  n % 8 == 0 */
int sfb_gelu_fwd(const void *x, void *y, int64_t n, void *stream);
int sfb_gelu_bwd(const void *dy, const void *x, void *dx, int64_t n, void *stream);

/* out[c][r] = in[r][c] for r < R, c < C (bf16); columns R <= r < ld_out of out are zero-filled, so ld_out can be R rounded up to the
 * multiple of 8 that sfb_gemm_bf16 needs as its K. */
int sfb_transpose_bf16(const void *in, int64_t ld_in, int R, int C, void *out, int64_t ld_out, void *stream);

/* out[c] = sum_r in[r][c]  (bias gradients; pos-emb / token gradients summed over the batch).  in is bf16 (in_bf16 != 0) or fp32,
 * N and ld even.  Deterministic two-stage reduction through `workspace` (up to 64 * N floats are used; NULL = single stage). */
int sfb_colsum(const void *in, int in_bf16, int64_t ld, int M, int N, float *out, float *workspace, int64_t workspace_floats, void *stream);

/* nn.LayerNorm backward over D = 768 in fp32 (transformer.py:84-85, sync_model.py:126-127,143):
 *     dx[r] (+)= rstd (g - mean(g) - xhat mean(g xhat)),  g = dy[row(r)] * gamma;   dgamma = sum_r dy xhat;   dbeta = sum_r dy
 * dy row of output row r: (r / group) * group_stride + offset + r % group (the gather of sfb_layernorm / the token layout of
 * sfb_sync_tokens).  accumulate != 0 adds into dx (residual branch).  dbeta must be dgamma + 768 (one (2, 768) buffer).
 * workspace: sfb_layernorm_bwd_workspace_floats(rows) floats. */
int sfb_layernorm_bwd_workspace_floats(int rows);
int sfb_layernorm_bwd(const float *dy, int64_t lddy, int group, int group_stride, int offset, const float *x, int64_t ldx,
                      const float *gamma, float eps, float *dx, int64_t lddx, int accumulate, float *dgamma, float *dbeta,
                      float *workspace, int64_t workspace_floats, int rows, void *stream);

/* SelfAttention.forward in training mode (transformer.py:58-76) on the fused qkv (B*T, 3*n_heads*head_dim) bf16 = [q | k | v]:
 *     out (B*T, n_heads*head_dim) bf16 = dropout(softmax(scale q k^T)) v,   lse (B, n_heads, T) fp32 (log2 units)
 * dropout element index e = ((b * n_heads + h) * T + i) * T + j.  head_dim in {64, 96}; T up to 487 (shared-memory bound). */
int sfb_attention_train_fwd(const void *qkv, void *out, float *lse, int B, int T, int n_heads, int head_dim, float scale, float p_drop,
                            uint64_t seed, uint32_t site, void *stream);
/* backward of the above: dqkv (B*T, 3*n_heads*head_dim) bf16 from d_out; P is recomputed from qkv and lse, the mask from the counter.
 * delta (B, n_heads, T) fp32 is scratch (dO . O per row). */
int sfb_attention_train_bwd(const void *qkv, const void *out, const void *d_out, const float *lse, float *delta, void *dqkv, int B, int T,
                            int n_heads, int head_dim, float scale, float p_drop, uint64_t seed, uint32_t site, void *stream);

/* backward of sfb_sync_head (sync_model.py:169-172): dx (B, T, 768) fp32 is zeroed and its token-0 rows receive the gradient;
 * dln_w / dln_b (768), dW (n_cls, 768), dbias (n_cls) fp32; scratch: B * 3 * 768 floats. */
int sfb_sync_head_bwd(const float *x, int T, const float *ln_w, const float *ln_b, float eps, const float *W, const float *dlogits, int B,
                      int n_cls, float *dx, float *dln_w, float *dln_b, float *dW, float *dbias, float *scratch, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * N3, optimiser side (scripts/train_utils.py:373-386: scaler.unscale_ -> clip_grad_norm_ -> Adam step; sync_model.py:91-99 loss).
 * A tensor list is described by a device table of n_tensors entries { float *param, const float *grad, float *exp_avg,
 * float *exp_avg_sq, int64_t n } (40 bytes each) plus, per chunk of sfb_optim_chunk_elems() elements, the index of its tensor
 * (chunk_tensor) and its first element (chunk_start).  All tensors fp32, contiguous. */
/* loss[0] = mean_b( logsumexp(logits[b]) - logits[b, targets[b]] ),  dlogits = (softmax(logits) - onehot(targets)) / B;  row_loss: B floats scratch */
int sfb_cross_entropy(const float *logits, const int64_t *targets, int B, int C, float *loss, float *dlogits, float *row_loss, void *stream);
int sfb_optim_chunk_elems(void);
/* sqnorm[0] = sum over all listed gradients of g^2 (partial: n_chunks floats scratch; fixed summation order) */
int sfb_grad_sqnorm(const void *table, const int32_t *chunk_tensor, const int64_t *chunk_start, int n_chunks, float *partial, float *sqnorm, void *stream);
/* torch.optim.Adam step (L2 weight decay, no amsgrad) on every listed tensor in one launch.  With norm = sqrt(sqnorm[0]) * inv_scale:
 * a non-finite norm skips the step and sets found_inf[0] = 1 (else 0); g = grad * inv_scale * min(1, max_norm / (norm + 1e-6))
 * (max_norm <= 0: no clipping).  step_count[0] (device, float) holds the number of completed steps and advances only if not skipped. */
int sfb_adam_step(const void *table, const int32_t *chunk_tensor, const int64_t *chunk_start, int n_chunks, const float *sqnorm, float *found_inf,
                  float *step_count, float lr, float beta1, float beta2, float eps, float weight_decay, float inv_scale, float max_norm, void *stream);

/* =========================================================================================================
 * N1 — backward of the encoders (stage-I contrastive training, open_clip/model.py:474-527, training/train.py:122-154).
 * Linear layers, LayerNorm, GELU, bias / embedding reductions and the AST attention (74 x 74, fused qkv) reuse the N3 entry points;
 * the entry points below add what the Motionformer's divided attention and the CLS aggregators need.
 * ========================================================================================================= */

/* Backward of sfb_attention(desc) for problems whose keys fit in shared memory (Lk + prefix <= ~480): time attention 8 x 9, space
 * attention 196 x 197 (vit_helper.py:126-146), CLS aggregators 1 x 197 / 1 x 13 (motionformer.py:301-334).  desc is the FORWARD
 * descriptor (desc->out = the forward output O; q_extra must be NULL); d_out has the strides of out, dq / dk / dv those of q / k / v.
 * The prefix key / value row is shared by the inner problems of an outer index: its gradient is written per problem to
 *     dprefix[((inner * n_outer + outer) * n_heads + head) * 2 * head_dim] = { dK[head_dim], dV[head_dim] }     (fp32)
 * so that one sfb_colsum over the n_inner (or all) problems reduces it deterministically.
 * stats: sfb_attention_bwd_stats_floats(desc) floats of scratch (row log-sum-exp and dO . O). */
int64_t sfb_attention_bwd_stats_floats(const sfb_attn_desc *desc);
int sfb_attention_bwd(const sfb_attn_desc *desc, const void *d_out, void *dq, void *dk, void *dv, float *dprefix, float *stats, void *stream);

/* Backward of the Motionformer CLS query (one query per (outer, head) over all Lk token rows of its outer index, vit_helper.py:124):
 * writes dq (1 row per outer), and ADDS this query's contribution to dk / dv of every token row - call it after sfb_attention_bwd has
 * written them.  prefix_grad (n_outer, n_heads, 2, head_dim) fp32, if given, is added to token row 0 (the reduced dprefix of the
 * divided problems: the CLS token's key / value).  coef: n_outer * n_heads * Lk * 2 floats of scratch.  head_dim 64. */
int sfb_attention_bwd_global_query(const void *q, int64_t q_outer, const void *k, const void *v, int64_t kv_outer, int64_t kv_row,
                                   const void *out, const void *d_out, int64_t o_outer, void *dq, void *dk, void *dv,
                                   const float *prefix_grad, float *coef, int n_outer, int n_heads, int head_dim, int Lk, float scale,
                                   void *stream);

/* timm DropPath (vit_helper.py:356,371,375; stochastic depth 0 .. 0.2 over the 12 blocks, divided_224_16x4.yaml:59):
 *     out = (residual ? residual : 0) + in * (keep(seed, site, sample) ? 1 / (1 - p) : 0),  sample = row / rows_per_sample
 * in / residual (n_rows, 768) fp32, out fp32 or bf16; forward and backward (same role as sfb_dropout). */
int sfb_droppath(const float *in, const float *residual, void *out, int out_bf16, int64_t n_rows, int rows_per_sample, float p,
                 uint64_t seed, uint32_t site, void *stream);

/* out[r] (bf16, 768 wide, contiguous) = in[(r / group) * group_stride + offset + r % group] (fp32, row stride ld): the row gather of
 * sfb_layernorm without the normalisation (patch-token gradients without the CLS rows, fp32 -> bf16 GEMM operands). */
int sfb_gather_rows_bf16(const float *in, int64_t ld, void *out, int rows, int group, int group_stride, int offset, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * N1, tail of the stage-I contrastive step: AVCLIP.forward after the two towers (open_clip/model.py:489-527, 536-545;
 * AveragePooling motionformer.py:405-409).  All tensors fp32, contiguous; deterministic (no atomics).
 * --------------------------------------------------------------------------------------------------------- */
/* out (n, D) = mean over the T token rows of x (n, T, D)  ['BS T D -> BS D'];  backward: dx (n, T, D) = dout / T */
int sfb_mean_tokens(const float *x, float *out, int n, int T, int D, void *stream);
int sfb_mean_tokens_bwd(const float *dout, float *dx, int n, int T, int D, void *stream);
/* F.normalize(x, dim=-1): xn = x / max(|x|, 1e-12); inv_norm (n) is kept for the backward dx = (dxn - xn (xn . dxn)) * inv_norm */
int sfb_l2_normalize(const float *x, float *xn, float *inv_norm, int n, int D, void *stream);
int sfb_l2_normalize_bwd(const float *xn, const float *inv_norm, const float *dxn, float *dx, int n, int D, void *stream);
/* compute_loss (open_clip/model.py:507-527): sim_v2a = vn an_all^T / scale, sim_a2v = an vn_all^T / scale with the n local rows as
 * queries and the N >= n gathered rows as keys (N = n without gather_for_loss), targets eye(n, N) exactly as the reference builds them,
 * loss[0] = (CE(sim_v2a) + CE(sim_a2v)) / 2.  scale: DEVICE pointer to the (clamped) logit_scale parameter - no host read.
 * Also written: dscale[0] = d loss / d scale, G (2, n, N) = d loss / d sim for the two directions (input of the backward).
 * row_ws: 4 n floats of scratch.  N <= 8192, D <= 4096. */
int sfb_contrastive_loss(const float *vn, const float *an, const float *vn_all, const float *an_all, int n, int N, int D, const float *scale,
                         float *loss, float *dscale, float *G, float *row_ws, void *stream);
/* gradients of the loss times the DEVICE scalar upstream[0] (the GradScaler's scale): local query rows d_vn / d_an (n, D), gathered key
 * rows d_vn_all / d_an_all (N, D) - reduce-scatter them over the ranks and add; pass NULL for both when N == n and the keys ARE the
 * queries (no gathering): their gradient is then added into d_vn / d_an.  dscale_out[0] = dscale[0] * upstream[0]. */
int sfb_contrastive_loss_bwd(const float *vn, const float *an, const float *vn_all, const float *an_all, const float *G, int n, int N, int D,
                             const float *scale, const float *upstream, const float *dscale, float *d_vn, float *d_an, float *d_vn_all,
                             float *d_an_all, float *dscale_out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SYNCHFORMER_B200_H */
