// K2 / K3 / K12 / K13: the data-movement kernels either side of the GEMMs — im2col gathers, token assembly,
// sync-transformer sequence assembly and the classification head.  All HBM-bound, 16-byte vector accesses,
// reads ordered so consecutive threads touch consecutive input bytes.
#include "common.cuh"

namespace sfb {

// ----------------------------------------------------------------------------------------------------------
// K2: PatchEmbed3D im2col (vit_helper.py:436-445).  One thread moves 8 consecutive pixels of one image row:
// idx -> (seg, t, c, Y, X8) is exactly the linear order of the input, so reads are perfectly coalesced; the
// 16-byte result lands at row seg*1568 + (t/2)*196 + (Y/16)*14 + X8/2, col c*512 + (t%2)*256 + (Y%16)*16 + (X8%2)*8.
// ----------------------------------------------------------------------------------------------------------
template <int DT>
__device__ __forceinline__ void load8(const void *base, int64_t idx, float *f) {
    if (DT == 0) {
        const float4 a = __ldg(reinterpret_cast<const float4 *>(base) + idx * 2);
        const float4 b = __ldg(reinterpret_cast<const float4 *>(base) + idx * 2 + 1);
        f[0] = a.x, f[1] = a.y, f[2] = a.z, f[3] = a.w, f[4] = b.x, f[5] = b.y, f[6] = b.z, f[7] = b.w;
    } else if (DT == 1) {
        const uint4 u = __ldg(reinterpret_cast<const uint4 *>(base) + idx);
        const __half2 *h = reinterpret_cast<const __half2 *>(&u);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 t = __half22float2(h[i]);
            f[2 * i] = t.x, f[2 * i + 1] = t.y;
        }
    } else if (DT == 2) {
        const uint4 u = __ldg(reinterpret_cast<const uint4 *>(base) + idx);
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 t = unpack_bf16x2(w[i]);
            f[2 * i] = t.x, f[2 * i + 1] = t.y;
        }
    } else {  // uint8 frames: /255 -> (x - 0.5) / 0.5 = 2x/255 - 1 with a single rounding   (dataset/transforms.py:647-669)
        const uint2 u = __ldg(reinterpret_cast<const uint2 *>(base) + idx);
        const uint32_t w[2] = {u.x, u.y};
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = fmaf(static_cast<float>((w[i >> 2] >> (8 * (i & 3))) & 0xffu), 2.0f / 255.0f, -1.0f);
    }
}

// n_segments == 0: `vis` already holds one 16-frame block per segment (idx is the input chunk index).  n_segments > 0 (N2, segment
// slicing fused into the gather, dataset/transforms.py:402-499): `vis` is (n_clips, n_frames, 3, 224, 224) and segment s of clip b
// covers frames [v_start + s * v_stride, + 16) - overlapping segments re-read shared frames from HBM/L2 instead of from the host.
template <int DT>
__global__ void __launch_bounds__(256) im2col_video_kernel(const void *__restrict__ vis, __nv_bfloat16 *__restrict__ A, int64_t n_chunks,
                                                           int n_frames, int n_segments, int v_start, int v_stride) {
    const int64_t idx = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
    if (idx >= n_chunks) return;
    const int X8 = static_cast<int>(idx % 28);
    int64_t r = idx / 28;
    const int Y = static_cast<int>(r % 224);
    r /= 224;
    const int c = static_cast<int>(r % 3);
    r /= 3;
    const int t = static_cast<int>(r % 16);
    const int64_t seg = r / 16;
    float f[8];
    int64_t in_idx = idx;
    if (n_segments > 0) {
        const int64_t frame = (seg / n_segments) * n_frames + v_start + (seg % n_segments) * v_stride + t;
        in_idx = ((frame * 3 + c) * 224 + Y) * 28 + X8;
    }
    load8<DT>(vis, in_idx, f);
    const int64_t row = seg * 1568 + (t >> 1) * 196 + (Y >> 4) * 14 + (X8 >> 1);
    const int col = c * 512 + (t & 1) * 256 + (Y & 15) * 16 + (X8 & 1) * 8;
    uint4 u;
    u.x = pack_bf16x2(f[0], f[1]), u.y = pack_bf16x2(f[2], f[3]), u.z = pack_bf16x2(f[4], f[5]), u.w = pack_bf16x2(f[6], f[7]);
    *reinterpret_cast<uint4 *>(A + row * 1536 + col) = u;
}

// video_model_builder.py:221-254: x[seg, 0] = cls + pos[0];  x[seg, 1 + f*196 + n] = patch + pos[1+n] + temp[f]
__global__ void __launch_bounds__(256) video_tokens_kernel(const float4 *__restrict__ patch, const float4 *__restrict__ cls,
                                                           const float4 *__restrict__ pos, const float4 *__restrict__ temp,
                                                           float4 *__restrict__ x, int64_t n_vec) {
    const int64_t idx = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
    if (idx >= n_vec) return;
    const int dv = static_cast<int>(idx % 192);
    const int64_t tokg = idx / 192;
    const int tok = static_cast<int>(tokg % 1569);
    const int64_t seg = tokg / 1569;
    float4 a, b;
    if (tok == 0) {
        a = __ldg(cls + dv), b = __ldg(pos + dv);
    } else {
        const int f = (tok - 1) / 196, n = (tok - 1) % 196;
        a = __ldg(patch + (seg * 1568 + tok - 1) * 192 + dv);
        const float4 p = __ldg(pos + (1 + n) * 192 + dv), t = __ldg(temp + f * 192 + dv);
        b = make_float4(p.x + t.x, p.y + t.y, p.z + t.z, p.w + t.w);
    }
    x[idx] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}

// ----------------------------------------------------------------------------------------------------------
// K3: AST patch embedding gather (modeling_ast.py:113-117): 16x16 patches, stride 10, of the (128 freq, 66 time) mel
// ----------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) im2col_ast_kernel(const float *__restrict__ spec, __nv_bfloat16 *__restrict__ A, int64_t n_chunks) {
    const int64_t idx = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
    if (idx >= n_chunks) return;
    const int j0 = static_cast<int>(idx & 1) * 8;
    const int i = static_cast<int>((idx >> 1) & 15);
    const int64_t row = idx >> 5;                 // seg*72 + f*6 + t
    const int t = static_cast<int>(row % 6), fq = static_cast<int>((row / 6) % 12);
    const int64_t seg = row / 72;
    const float *sp = spec + (seg * 128 + fq * 10 + i) * 66 + t * 10 + j0;
    uint4 u;
    u.x = pack_bf16x2(__ldg(sp), __ldg(sp + 1)), u.y = pack_bf16x2(__ldg(sp + 2), __ldg(sp + 3));
    u.z = pack_bf16x2(__ldg(sp + 4), __ldg(sp + 5)), u.w = pack_bf16x2(__ldg(sp + 6), __ldg(sp + 7));
    *reinterpret_cast<uint4 *>(A + row * 256 + i * 16 + j0) = u;
}

// modeling_ast.py:83-93: [cls, distillation, patches] + position_embeddings
__global__ void __launch_bounds__(256) ast_tokens_kernel(const float4 *__restrict__ patch, const float4 *__restrict__ cls,
                                                         const float4 *__restrict__ dist, const float4 *__restrict__ pos,
                                                         float4 *__restrict__ x, int64_t n_vec) {
    const int64_t idx = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
    if (idx >= n_vec) return;
    const int dv = static_cast<int>(idx % 192);
    const int64_t tokg = idx / 192;
    const int tok = static_cast<int>(tokg % 74);
    const int64_t seg = tokg / 74;
    const float4 a = tok == 0 ? __ldg(cls + dv) : tok == 1 ? __ldg(dist + dv) : __ldg(patch + (seg * 72 + tok - 2) * 192 + dv);
    const float4 b = __ldg(pos + tok * 192 + dv);
    x[idx] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}

// ----------------------------------------------------------------------------------------------------------
// K12: sync_model.py:150-167 — one warp per token of the (B, 2+14S, 768) sequence
// ----------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sync_tokens_kernel(const float *__restrict__ v, const float *__restrict__ a,
                                                          const float *__restrict__ vg, const float *__restrict__ vb,
                                                          const float *__restrict__ ag, const float *__restrict__ ab, float eps,
                                                          const float *__restrict__ off_tok, const float *__restrict__ mod_tok,
                                                          const float *__restrict__ pos, float *__restrict__ x, int B, int S) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int T = 2 + 14 * S;
    const int64_t r = static_cast<int64_t>(blockIdx.x) * 8 + warp;
    if (r >= static_cast<int64_t>(B) * T) return;
    const int tok = static_cast<int>(r % T);
    const int64_t b = r / T;
    const float4 *pp = reinterpret_cast<const float4 *>(pos + static_cast<int64_t>(tok) * kD);
    float4 *xp = reinterpret_cast<float4 *>(x + r * kD);
    float4 val[6];
    if (tok == 0 || tok == 8 * S + 1) {
        const float4 *tp = reinterpret_cast<const float4 *>(tok == 0 ? off_tok : mod_tok);
#pragma unroll
        for (int j = 0; j < 6; ++j) val[j] = __ldg(tp + lane + 32 * j);
    } else {
        const bool is_v = tok <= 8 * S;
        const float *src = is_v ? v + (b * 8 * S + (tok - 1)) * kD : a + (b * 6 * S + (tok - 8 * S - 2)) * kD;
        const float4 *g = reinterpret_cast<const float4 *>(is_v ? vg : ag), *bt = reinterpret_cast<const float4 *>(is_v ? vb : ab);
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            val[j] = __ldg(reinterpret_cast<const float4 *>(src) + lane + 32 * j);
            s += (val[j].x + val[j].y) + (val[j].z + val[j].w);
        }
        const float mean = warp_sum(s) * (1.0f / kD);
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            val[j].x -= mean, val[j].y -= mean, val[j].z -= mean, val[j].w -= mean;
            q += (val[j].x * val[j].x + val[j].y * val[j].y) + (val[j].z * val[j].z + val[j].w * val[j].w);
        }
        const float rstd = rsqrtf(warp_sum(q) * (1.0f / kD) + eps);
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            const float4 gg = __ldg(g + lane + 32 * j), bb = __ldg(bt + lane + 32 * j);
            val[j] = make_float4(val[j].x * rstd * gg.x + bb.x, val[j].y * rstd * gg.y + bb.y, val[j].z * rstd * gg.z + bb.z,
                                 val[j].w * rstd * gg.w + bb.w);
        }
    }
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        const float4 p = __ldg(pp + lane + 32 * j);
        xp[lane + 32 * j] = make_float4(val[j].x + p.x, val[j].y + p.y, val[j].z + p.z, val[j].w + p.w);
    }
}

// K13: sync_model.py:169-172 — ln_f on token 0 then Linear(768 -> n_cls), fp32 throughout; one CTA per clip
__global__ void __launch_bounds__(256) sync_head_kernel(const float *__restrict__ x, int T, const float *__restrict__ g,
                                                        const float *__restrict__ bt, float eps, const float *__restrict__ W,
                                                        const float *__restrict__ bias, float *__restrict__ logits, int n_cls) {
    __shared__ float xn[kD];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float *row = x + static_cast<int64_t>(blockIdx.x) * T * kD;
    if (warp == 0) {
        float val[24], s = 0.f;
#pragma unroll
        for (int j = 0; j < 24; ++j) val[j] = row[lane + 32 * j], s += val[j];
        const float mean = warp_sum(s) * (1.0f / kD);
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < 24; ++j) val[j] -= mean, q += val[j] * val[j];
        const float rstd = rsqrtf(warp_sum(q) * (1.0f / kD) + eps);
#pragma unroll
        for (int j = 0; j < 24; ++j) xn[lane + 32 * j] = val[j] * rstd * g[lane + 32 * j] + bt[lane + 32 * j];
    }
    __syncthreads();
    for (int c = warp; c < n_cls; c += 8) {
        float s = 0.f;
        for (int j = lane; j < kD; j += 32) s = fmaf(xn[j], W[static_cast<int64_t>(c) * kD + j], s);
        s = warp_sum(s);
        if (lane == 0) logits[static_cast<int64_t>(blockIdx.x) * n_cls + c] = s + bias[c];
    }
}

__global__ void __launch_bounds__(256) cast_f32_bf16_kernel(const float4 *__restrict__ in, uint2 *__restrict__ out, int64_t n_vec) {
    const int64_t idx = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
    if (idx >= n_vec) return;
    const float4 v = __ldg(in + idx);
    out[idx] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
}

static inline unsigned blocks_for(int64_t n) { return static_cast<unsigned>((n + 255) / 256); }
static inline bool al16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace sfb

static int im2col_video_launch(const void *vis, int in_dtype, void *A, int64_t n_seg, int n_frames, int n_segments, int v_start, int v_stride,
                               cudaStream_t st) {
    using namespace sfb;
    const int64_t n_chunks = n_seg * 16 * 3 * 224 * 28;
    SFB_CHECK_ARG(blocks_for(n_chunks) < (1u << 31), "sfb_im2col_video: too many segments for one launch");
    __nv_bfloat16 *a = reinterpret_cast<__nv_bfloat16 *>(A);
    switch (in_dtype) {
        case 0: im2col_video_kernel<0><<<blocks_for(n_chunks), 256, 0, st>>>(vis, a, n_chunks, n_frames, n_segments, v_start, v_stride); break;
        case 1: im2col_video_kernel<1><<<blocks_for(n_chunks), 256, 0, st>>>(vis, a, n_chunks, n_frames, n_segments, v_start, v_stride); break;
        case 2: im2col_video_kernel<2><<<blocks_for(n_chunks), 256, 0, st>>>(vis, a, n_chunks, n_frames, n_segments, v_start, v_stride); break;
        case 3: im2col_video_kernel<3><<<blocks_for(n_chunks), 256, 0, st>>>(vis, a, n_chunks, n_frames, n_segments, v_start, v_stride); break;
        default: set_error("sfb_im2col_video: in_dtype %d not in {0 f32, 1 f16, 2 bf16, 3 u8}", in_dtype); return SFB_E_INVALID;
    }
    SFB_CHECK_LAUNCH();
    return SFB_OK;
}

extern "C" int sfb_im2col_video(const void *vis, int in_dtype, void *A, int n_seg, void *stream) {
    using namespace sfb;
    SFB_CHECK_ARG(vis && A && n_seg > 0, "sfb_im2col_video: bad arguments");
    SFB_CHECK_ARG(al16(vis) && al16(A), "sfb_im2col_video: pointers must be 16-byte aligned");
    return im2col_video_launch(vis, in_dtype, A, n_seg, 0, 0, 0, 0, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int sfb_im2col_video_clip(const void *clip, int in_dtype, void *A, int n_clips, int n_frames, int n_segments, int v_start,
                                     int v_stride, void *stream) {
    using namespace sfb;
    SFB_CHECK_ARG(clip && A && n_clips > 0 && n_frames > 0 && n_segments > 0 && v_start >= 0 && v_stride > 0, "sfb_im2col_video_clip: bad arguments");
    SFB_CHECK_ARG(v_start + (n_segments - 1) * v_stride + 16 <= n_frames, "sfb_im2col_video_clip: %d segments of 16 frames (start %d, stride %d) do not fit in %d frames",
                  n_segments, v_start, v_stride, n_frames);
    SFB_CHECK_ARG(al16(clip) && al16(A), "sfb_im2col_video_clip: pointers must be 16-byte aligned");
    return im2col_video_launch(clip, in_dtype, A, static_cast<int64_t>(n_clips) * n_segments, n_frames, n_segments, v_start, v_stride,
                               reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int sfb_video_tokens(const float *patch, const float *cls_token, const float *pos_embed, const float *temp_embed, float *x,
                                int n_seg, void *stream) {
    using namespace sfb;
    SFB_CHECK_ARG(patch && cls_token && pos_embed && temp_embed && x && n_seg > 0, "sfb_video_tokens: bad arguments");
    SFB_CHECK_ARG(al16(patch) && al16(cls_token) && al16(pos_embed) && al16(temp_embed) && al16(x), "sfb_video_tokens: alignment");
    const int64_t n_vec = static_cast<int64_t>(n_seg) * 1569 * 192;
    video_tokens_kernel<<<blocks_for(n_vec), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float4 *>(patch), reinterpret_cast<const float4 *>(cls_token), reinterpret_cast<const float4 *>(pos_embed),
        reinterpret_cast<const float4 *>(temp_embed), reinterpret_cast<float4 *>(x), n_vec);
    SFB_CHECK_LAUNCH();
    return SFB_OK;
}

extern "C" int sfb_im2col_ast(const float *spec, void *A, int n_seg, void *stream) {
    using namespace sfb;
    SFB_CHECK_ARG(spec && A && n_seg > 0 && al16(A), "sfb_im2col_ast: bad arguments");
    const int64_t n_chunks = static_cast<int64_t>(n_seg) * 72 * 32;
    im2col_ast_kernel<<<blocks_for(n_chunks), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(spec, reinterpret_cast<__nv_bfloat16 *>(A), n_chunks);
    SFB_CHECK_LAUNCH();
    return SFB_OK;
}

extern "C" int sfb_ast_tokens(const float *patch, const float *cls_token, const float *dist_token, const float *pos_embed, float *x,
                              int n_seg, void *stream) {
    using namespace sfb;
    SFB_CHECK_ARG(patch && cls_token && dist_token && pos_embed && x && n_seg > 0, "sfb_ast_tokens: bad arguments");
    SFB_CHECK_ARG(al16(patch) && al16(cls_token) && al16(dist_token) && al16(pos_embed) && al16(x), "sfb_ast_tokens: alignment");
    const int64_t n_vec = static_cast<int64_t>(n_seg) * 74 * 192;
    ast_tokens_kernel<<<blocks_for(n_vec), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float4 *>(patch), reinterpret_cast<const float4 *>(cls_token), reinterpret_cast<const float4 *>(dist_token),
        reinterpret_cast<const float4 *>(pos_embed), reinterpret_cast<float4 *>(x), n_vec);
    SFB_CHECK_LAUNCH();
    return SFB_OK;
}

extern "C" int sfb_sync_tokens(const float *v, const float *a, const float *vis_ln_w, const float *vis_ln_b, const float *aud_ln_w,
                               const float *aud_ln_b, float eps, const float *off_tok, const float *mod_tok, const float *pos_emb,
                               float *x, int B, int S, void *stream) {
    using namespace sfb;
    SFB_CHECK_ARG(v && a && vis_ln_w && vis_ln_b && aud_ln_w && aud_ln_b && off_tok && mod_tok && pos_emb && x, "sfb_sync_tokens: null pointer");
    SFB_CHECK_ARG(B > 0 && S > 0, "sfb_sync_tokens: B=%d S=%d", B, S);
    SFB_CHECK_ARG(al16(v) && al16(a) && al16(x) && al16(pos_emb) && al16(off_tok) && al16(mod_tok) && al16(vis_ln_w) && al16(vis_ln_b) &&
                      al16(aud_ln_w) && al16(aud_ln_b),
                  "sfb_sync_tokens: alignment");
    const int64_t rows = static_cast<int64_t>(B) * (2 + 14 * S);
    sync_tokens_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        v, a, vis_ln_w, vis_ln_b, aud_ln_w, aud_ln_b, eps, off_tok, mod_tok, pos_emb, x, B, S);
    SFB_CHECK_LAUNCH();
    return SFB_OK;
}

extern "C" int sfb_sync_head(const float *x, int T, const float *ln_w, const float *ln_b, float eps, const float *W, const float *b,
                             float *logits, int B, int n_cls, void *stream) {
    using namespace sfb;
    SFB_CHECK_ARG(x && ln_w && ln_b && W && b && logits && B > 0 && T > 0 && n_cls > 0, "sfb_sync_head: bad arguments");
    sync_head_kernel<<<B, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, T, ln_w, ln_b, eps, W, b, logits, n_cls);
    SFB_CHECK_LAUNCH();
    return SFB_OK;
}

extern "C" int sfb_cast_f32_bf16(const float *in, void *out, int64_t n, void *stream) {
    using namespace sfb;
    SFB_CHECK_ARG(in && out && n > 0 && n % 4 == 0 && al16(in) && al16(out), "sfb_cast_f32_bf16: bad arguments");
    cast_f32_bf16_kernel<<<blocks_for(n / 4), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const float4 *>(in),
                                                                                              reinterpret_cast<uint2 *>(out), n / 4);
    SFB_CHECK_LAUNCH();
    return SFB_OK;
}
