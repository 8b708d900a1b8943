// K1: mel front-end on the GPU (the reference runs it in CPU dataloader workers through torchaudio).
//   dataset/transforms.py:815-823  MelSpectrogram(sr 16 kHz, n_fft 1024, win 400, hop 160, 128 mels, power 2)
//   dataset/transforms.py:826-834  log(x + 1e-6)
//   dataset/transforms.py:836-858  pad time 65 -> 66 with 0.0 (log domain, before normalisation)
//   dataset/transforms.py:861-871  (x - (-4.2677393)) / (2 * 4.5689974)
// One CTA (256 threads) per (segment, STFT frame).  Only the 400 samples under the periodic Hann window are non-zero in the
// 1024-sample frame, and |X|^2 does not depend on where they sit, so they are placed at positions 0..399.  The real 1024-point
// transform is ONE 512-point complex FFT of z[n] = x[2n] + i x[2n+1] (radix-2 decimation in frequency in shared memory: 9 stages of 256
// butterflies, natural-order input, bit-reversed output) followed by the even / odd split
//     X[k] = E[k] + W^k O[k],   X[512 - k] = conj(E[k] - W^k O[k]),   E = (Z[k] + conj Z[512-k]) / 2,   O = (Z[k] - conj Z[512-k]) / 2i
// - 2 304 butterflies per frame where the round-1 kernel spent 103 000 multiply-adds on a direct 400-term DFT per bin pair.
// Everything up to |X|^2 stays in fp64: a pure tone leaves most bins ~1e-8 of the peak and log(x + 1e-6) exposes fp32 round-off there
// (torchaudio's fp32 FFT sits 2.3e-4 from the fp64 oracle for that reason).
// The 513 x 128 HTK triangle filterbank (L2-resident, 262 KB) is applied from shared-memory power values.
#include <math.h>

#include "common.cuh"

namespace sfb {
namespace mel {

constexpr int N_FFT = 1024, WIN = 400, HOP = 160, N_FREQ = 513, N_MEL = 128, N_FRAMES = 65, T_OUT = 66, SEG = 10240;
constexpr int WIN_OFF = (N_FFT - WIN) / 2;  // 312
constexpr float LOG_EPS = 1e-6f, NORM_MEAN = -4.2677393f, NORM_STD = 4.5689974f;

__device__ double2 g_twiddle[N_FFT];       // (cos, sin)(2 pi j / 1024)
__device__ float g_window[WIN];
__device__ float g_fb[N_FREQ * N_MEL];     // [freq][mel]

constexpr int MEL_THREADS = 256;
constexpr int N_HALF = N_FFT / 2;          // length of the complex transform

__global__ void __launch_bounds__(MEL_THREADS) mel_kernel(const float *__restrict__ wave, float *__restrict__ out, int n_segments,
                                                          int64_t clip_stride, int a_start, int a_stride) {
    __shared__ double2 z[N_HALF];
    __shared__ float pw[N_FREQ + 3];
    const int frame = blockIdx.x % N_FRAMES;
    const int64_t seg = blockIdx.x / N_FRAMES;
    const int tid = threadIdx.x;
    // segment `seg` = segment (seg % n_segments) of clip (seg / n_segments): windows may overlap inside one un-duplicated waveform
    const float *w = wave + (seg / n_segments) * clip_stride + a_start + (seg % n_segments) * static_cast<int64_t>(a_stride);
    auto sample = [&](int n) -> double {
        int idx = frame * HOP + WIN_OFF + n - N_FFT / 2;          // index into the un-padded segment
        idx = idx < 0 ? -idx : (idx >= SEG ? 2 * (SEG - 1) - idx : idx);   // reflect padding (torch.stft center=True)
        return static_cast<double>(__ldg(w + idx) * g_window[n]);          // windowed in fp32, exactly like the oracle / torchaudio
    };
    for (int n = tid; n < N_HALF; n += MEL_THREADS) z[n] = 2 * n < WIN ? make_double2(sample(2 * n), sample(2 * n + 1)) : make_double2(0.0, 0.0);
    __syncthreads();
#pragma unroll 1
    for (int half = N_HALF / 2; half >= 1; half >>= 1) {            // stage with butterflies of span `half`: one per thread
        const int pos = tid & (half - 1);
        const int i0 = ((tid - pos) << 1) + pos, i1 = i0 + half;
        const double2 a = z[i0], b = z[i1];
        const double2 t = g_twiddle[pos * (N_HALF / half)];          // (cos, sin)(2 pi pos / (2 half))
        const double dr = a.x - b.x, di = a.y - b.y;
        z[i0] = make_double2(a.x + b.x, a.y + b.y);
        z[i1] = make_double2(fma(dr, t.x, di * t.y), fma(di, t.x, -dr * t.y));     // (dr + i di) e^{-i theta}
        __syncthreads();
    }
    for (int k = tid; k <= N_HALF / 2; k += MEL_THREADS) {           // bin pairs (k, 512 - k); Z[k] sits at the bit-reversed index
        const double2 zk = z[__brev(static_cast<unsigned>(k)) >> 23];
        const double2 zm = z[__brev(static_cast<unsigned>((N_HALF - k) & (N_HALF - 1))) >> 23];
        const double er = 0.5 * (zk.x + zm.x), ei = 0.5 * (zk.y - zm.y);          // E = (Z[k] + conj Z[512-k]) / 2
        const double orr = 0.5 * (zk.y + zm.y), oi = -0.5 * (zk.x - zm.x);        // O = (Z[k] - conj Z[512-k]) / 2i
        const double2 t = g_twiddle[k];                                           // W^k = cos - i sin
        const double wr = fma(orr, t.x, oi * t.y), wi = fma(oi, t.x, -orr * t.y);
        const double r0 = er + wr, i0 = ei + wi, r1 = er - wr, i1 = ei - wi;
        pw[k] = static_cast<float>(r0 * r0 + i0 * i0);
        if (k != N_HALF - k) pw[N_HALF - k] = static_cast<float>(r1 * r1 + i1 * i1);
    }
    __syncthreads();
    if (tid < N_MEL) {
        float acc = 0.f;
        for (int k = 0; k < N_FREQ; ++k) acc = fmaf(pw[k], __ldg(g_fb + k * N_MEL + tid), acc);
        float *o = out + (seg * N_MEL + tid) * T_OUT;
        o[frame] = (logf(acc + LOG_EPS) - NORM_MEAN) / (2.0f * NORM_STD);
        if (frame == 0) o[T_OUT - 1] = (0.0f - NORM_MEAN) / (2.0f * NORM_STD);
    }
}

static int init_tables() {
    static PerDeviceOnce once;            // the tables live in per-device global memory: upload them on every device that is used
    if (!once.first()) return SFB_OK;
    static double2 tw[N_FFT];
    static float win[WIN];
    static float fb[N_FREQ * N_MEL];
    const double pi = 3.14159265358979323846;
    for (int j = 0; j < N_FFT; ++j) tw[j] = make_double2(cos(2.0 * pi * j / N_FFT), sin(2.0 * pi * j / N_FFT));
    for (int n = 0; n < WIN; ++n) win[n] = static_cast<float>(0.5 - 0.5 * cos(2.0 * pi * n / WIN));   // periodic Hann
    // torchaudio.functional.melscale_fbanks(513, 0, 8000, 128, 16000, norm=None, mel_scale='htk')
    auto hz2mel = [](double f) { return 2595.0 * log10(1.0 + f / 700.0); };
    double fpts[N_MEL + 2];
    const double m_lo = hz2mel(0.0), m_hi = hz2mel(8000.0);
    for (int i = 0; i < N_MEL + 2; ++i) {
        const double m = m_lo + (m_hi - m_lo) * i / (N_MEL + 1);
        fpts[i] = 700.0 * (pow(10.0, m / 2595.0) - 1.0);
    }
    for (int k = 0; k < N_FREQ; ++k) {
        const double f = 8000.0 * k / (N_FREQ - 1);
        for (int m = 0; m < N_MEL; ++m) {
            const double down = (f - fpts[m]) / (fpts[m + 1] - fpts[m]);
            const double up = (fpts[m + 2] - f) / (fpts[m + 2] - fpts[m + 1]);
            const double v = fmin(down, up);
            fb[k * N_MEL + m] = static_cast<float>(v > 0.0 ? v : 0.0);
        }
    }
    SFB_CHECK_CUDA(cudaMemcpyToSymbol(g_twiddle, tw, sizeof(tw)));
    SFB_CHECK_CUDA(cudaMemcpyToSymbol(g_window, win, sizeof(win)));
    SFB_CHECK_CUDA(cudaMemcpyToSymbol(g_fb, fb, sizeof(fb)));
    return SFB_OK;
}

}  // namespace mel
}  // namespace sfb

extern "C" int sfb_mel_frontend(const float *wave, float *out, int n_seg, void *stream) {
    using namespace sfb;
    using namespace sfb::mel;
    SFB_CHECK_ARG(wave && out && n_seg > 0, "sfb_mel_frontend: bad arguments");
    int rc = init_tables();   // first call only: three small host->device table uploads
    if (rc != SFB_OK) return rc;
    mel_kernel<<<static_cast<unsigned>(static_cast<int64_t>(n_seg) * N_FRAMES), MEL_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(wave, out, 1, SEG, 0, 0);
    SFB_CHECK_LAUNCH();
    return SFB_OK;
}

extern "C" int sfb_mel_frontend_clip(const float *wave, int64_t clip_stride, float *out, int n_clips, int n_segments, int a_start, int a_stride,
                                     void *stream) {
    using namespace sfb;
    using namespace sfb::mel;
    SFB_CHECK_ARG(wave && out && n_clips > 0 && n_segments > 0 && a_start >= 0 && a_stride > 0, "sfb_mel_frontend_clip: bad arguments");
    SFB_CHECK_ARG(a_start + static_cast<int64_t>(n_segments - 1) * a_stride + SEG <= clip_stride, "sfb_mel_frontend_clip: segments do not fit in the waveform");
    int rc = init_tables();
    if (rc != SFB_OK) return rc;
    mel_kernel<<<static_cast<unsigned>(static_cast<int64_t>(n_clips) * n_segments * N_FRAMES), MEL_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        wave, out, n_segments, clip_stride, a_start, a_stride);
    SFB_CHECK_LAUNCH();
    return SFB_OK;
}
