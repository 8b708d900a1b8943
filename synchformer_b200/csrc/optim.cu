// N3, optimiser side of the training step (scripts/train_utils.py:373-386: GradScaler unscale -> clip_grad_norm_ -> Adam):
//
//   sfb_cross_entropy        F.cross_entropy(logits, targets) (mean) fused with its gradient (compute_loss, sync_model.py:91-99)
//   sfb_grad_sqnorm          sum of squares over a LIST of gradient tensors (one launch, deterministic two-stage reduction)
//   sfb_adam_step            torch.optim.Adam (L2 weight decay, no amsgrad) over the same list in ONE launch, with the GradScaler's
//                            1 / scale, the clip_grad_norm_ coefficient min(1, max_norm / (norm + 1e-6)) and the skip-on-inf/nan
//                            rule all read from the device-side norm: the host never synchronises
//
// The tensor list is a device table: per tensor { param, grad, exp_avg, exp_avg_sq pointers, element count }, and per 64 K-element
// chunk the tensor it belongs to and its start offset.  22.6 M parameters = 63 tensors = 380 chunks: HBM-bound
// (16 B read + 12 B written per element), one pass.
#include "common.cuh"

namespace sfb {
namespace optim {

constexpr int kChunk = 65536;

struct TensorEntry {
    float *param;
    const float *grad;
    float *exp_avg;
    float *exp_avg_sq;
    int64_t n;
};

// ---------------------------------------------------------------------------------------------------------------------------
// cross entropy: warp per row; row_loss[b] = logsumexp(logits[b]) - logits[b, target[b]];  dlogits = (softmax - onehot) / n_valid
// Targets follow F.cross_entropy: -100 (its default ignore_index) drops the row from the loss, the gradient and the mean's denominator;
// any other value outside [0, C) is an error - torch raises a device assert, here the loss and that row's gradient become NaN (loud,
// without killing the context) and nothing is read out of bounds.
// ---------------------------------------------------------------------------------------------------------------------------
constexpr int64_t kIgnoreIndex = -100;

__device__ __forceinline__ int count_valid_targets(const int64_t *__restrict__ targets, int B, int lane) {
    int n = 0;
    for (int b = lane; b < B; b += 32) n += targets[b] != kIgnoreIndex ? 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
    return n;
}

__global__ void __launch_bounds__(256) cross_entropy_rows_kernel(const float *__restrict__ logits, const int64_t *__restrict__ targets, int B, int C,
                                                                 float *__restrict__ row_loss, float *__restrict__ dlogits) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * 8 + warp;
    if (b >= B) return;
    const int n_valid = count_valid_targets(targets, B, lane);
    const float *row = logits + static_cast<int64_t>(b) * C;
    const int64_t t64 = targets[b];
    if (t64 == kIgnoreIndex) {
        for (int c = lane; c < C; c += 32) dlogits[static_cast<int64_t>(b) * C + c] = 0.0f;
        if (lane == 0) row_loss[b] = 0.0f;
        return;
    }
    const bool bad = t64 < 0 || t64 >= C;
    float mx = -INFINITY;
    for (int c = lane; c < C; c += 32) mx = fmaxf(mx, row[c]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int c = lane; c < C; c += 32) sum += expf(row[c] - mx);
    sum = warp_sum(sum);
    const int t = bad ? 0 : static_cast<int>(t64);
    const float inv = 1.0f / sum, invB = 1.0f / n_valid, poison = bad ? __uint_as_float(0x7fc00000u) : 0.0f;
    for (int c = lane; c < C; c += 32) dlogits[static_cast<int64_t>(b) * C + c] = (expf(row[c] - mx) * inv - (c == t ? 1.0f : 0.0f)) * invB + poison;
    if (lane == 0) row_loss[b] = mx + logf(sum) - row[t] + poison;
}

// out = sum(v) / (number of targets that are not ignore_index); all rows ignored -> NaN, as F.cross_entropy
__global__ void __launch_bounds__(256) mean_kernel(const float *__restrict__ v, const int64_t *__restrict__ targets, int n, float *__restrict__ out) {
    __shared__ float red[8];
    float s = 0.f;
    for (int i = threadIdx.x; i < n; i += 256) s += v[i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        const int n_valid = count_valid_targets(targets, n, threadIdx.x);
        if (threadIdx.x == 0) {
            float t = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) t += red[k];
            out[0] = t / n_valid;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// sum of squares of all gradients: CTA per chunk -> partial[chunk]; then one CTA adds the partials in a fixed order
// ---------------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sqnorm_chunks_kernel(const TensorEntry *__restrict__ table, const int32_t *__restrict__ chunk_tensor,
                                                            const int64_t *__restrict__ chunk_start, float *__restrict__ partial) {
    __shared__ float red[8];
    const TensorEntry e = table[chunk_tensor[blockIdx.x]];
    const int64_t start = chunk_start[blockIdx.x];
    const int64_t end = start + kChunk < e.n ? start + kChunk : e.n;
    float s = 0.f;
    for (int64_t i = start + threadIdx.x; i < end; i += 256) {
        const float g = e.grad[i];
        s = fmaf(g, g, s);
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += red[k];
        partial[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(256) sum_kernel(const float *__restrict__ v, int n, float *__restrict__ out) {
    __shared__ float red[8];
    float s = 0.f;
    for (int i = threadIdx.x; i < n; i += 256) s += v[i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += red[k];
        out[0] = t;
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// Adam over all chunks.  sqnorm = sum of squares of the SCALED gradients (device scalar), inv_scale = 1 / GradScaler scale.
//   norm = sqrt(sqnorm) * inv_scale;  non-finite -> the whole step is skipped (found_inf[0] = 1), as GradScaler.step does
//   g = grad * inv_scale * min(1, max_norm / (norm + 1e-6)) + weight_decay * p          (max_norm <= 0: no clipping)
//   m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g g;  p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps),  t = step_count + 1
// ---------------------------------------------------------------------------------------------------------------------------
struct AdamParams {
    float lr, beta1, beta2, eps, weight_decay, inv_scale, max_norm;
};

__global__ void __launch_bounds__(256) adam_chunks_kernel(const TensorEntry *__restrict__ table, const int32_t *__restrict__ chunk_tensor,
                                                          const int64_t *__restrict__ chunk_start, const float *__restrict__ sqnorm,
                                                          float *__restrict__ found_inf, const float *__restrict__ step_count, AdamParams a) {
    const float norm = sqrtf(sqnorm[0]) * a.inv_scale;
    if (!isfinite(norm)) {
        if (blockIdx.x == 0 && threadIdx.x == 0) found_inf[0] = 1.0f;
        return;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) found_inf[0] = 0.0f;
    float coef = a.inv_scale;
    if (a.max_norm > 0.f) coef *= fminf(1.0f, a.max_norm / (norm + 1e-6f));
    const TensorEntry e = table[chunk_tensor[blockIdx.x]];
    const int64_t start = chunk_start[blockIdx.x];
    const int64_t end = start + kChunk < e.n ? start + kChunk : e.n;
    const float t = step_count[0] + 1.0f;                 // completed steps live on the device: a skipped step must not advance them
    const float step_size = a.lr / (1.0f - powf(a.beta1, t));
    const float bc2_sqrt = sqrtf(1.0f - powf(a.beta2, t));
    for (int64_t i = start + threadIdx.x; i < end; i += 256) {
        const float p = e.param[i];
        const float g = fmaf(a.weight_decay, p, e.grad[i] * coef);
        const float m = a.beta1 * e.exp_avg[i] + (1.0f - a.beta1) * g;
        const float v = a.beta2 * e.exp_avg_sq[i] + (1.0f - a.beta2) * g * g;
        e.exp_avg[i] = m;
        e.exp_avg_sq[i] = v;
        e.param[i] = p - step_size * m / (sqrtf(v) / bc2_sqrt + a.eps);
    }
}

// step_count += 1 unless the step was skipped; runs after adam_chunks_kernel on the same stream
__global__ void bump_step_kernel(const float *__restrict__ found_inf, float *__restrict__ step_count) {
    if (threadIdx.x == 0 && blockIdx.x == 0 && found_inf[0] == 0.0f) step_count[0] += 1.0f;
}

}  // namespace optim
}  // namespace sfb

extern "C" int sfb_cross_entropy(const float *logits, const int64_t *targets, int B, int C, float *loss, float *dlogits, float *row_loss, void *stream) {
    using namespace sfb;
    using namespace sfb::optim;
    SFB_CHECK_ARG(logits && targets && loss && dlogits && row_loss && B > 0 && C > 0, "sfb_cross_entropy: bad arguments");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    cross_entropy_rows_kernel<<<(B + 7) / 8, 256, 0, st>>>(logits, targets, B, C, row_loss, dlogits);
    SFB_CHECK_LAUNCH();
    mean_kernel<<<1, 256, 0, st>>>(row_loss, targets, B, loss);
    SFB_CHECK_LAUNCH();
    return SFB_OK;
}

extern "C" int sfb_optim_chunk_elems(void) { return sfb::optim::kChunk; }

extern "C" int sfb_grad_sqnorm(const void *table, const int32_t *chunk_tensor, const int64_t *chunk_start, int n_chunks, float *partial, float *sqnorm,
                               void *stream) {
    using namespace sfb;
    using namespace sfb::optim;
    SFB_CHECK_ARG(table && chunk_tensor && chunk_start && partial && sqnorm && n_chunks > 0, "sfb_grad_sqnorm: bad arguments");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    sqnorm_chunks_kernel<<<n_chunks, 256, 0, st>>>(reinterpret_cast<const TensorEntry *>(table), chunk_tensor, chunk_start, partial);
    SFB_CHECK_LAUNCH();
    sum_kernel<<<1, 256, 0, st>>>(partial, n_chunks, sqnorm);
    SFB_CHECK_LAUNCH();
    return SFB_OK;
}

extern "C" int sfb_adam_step(const void *table, const int32_t *chunk_tensor, const int64_t *chunk_start, int n_chunks, const float *sqnorm, float *found_inf,
                             float *step_count, float lr, float beta1, float beta2, float eps, float weight_decay, float inv_scale, float max_norm,
                             void *stream) {
    using namespace sfb;
    using namespace sfb::optim;
    SFB_CHECK_ARG(table && chunk_tensor && chunk_start && sqnorm && found_inf && step_count && n_chunks > 0, "sfb_adam_step: bad arguments");
    SFB_CHECK_ARG(beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && eps >= 0.f && lr >= 0.f, "sfb_adam_step: bad hyper-parameters");
    AdamParams a;
    a.lr = lr, a.beta1 = beta1, a.beta2 = beta2, a.eps = eps, a.weight_decay = weight_decay, a.inv_scale = inv_scale, a.max_norm = max_norm;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    adam_chunks_kernel<<<n_chunks, 256, 0, st>>>(reinterpret_cast<const TensorEntry *>(table), chunk_tensor, chunk_start, sqnorm, found_inf, step_count, a);
    SFB_CHECK_LAUNCH();
    bump_step_kernel<<<1, 32, 0, st>>>(found_inf, step_count);
    SFB_CHECK_LAUNCH();
    return SFB_OK;
}
