// K1 / K2 front end:  PCM segment -> log-mel dB -> standardised 512x512 image -> stem im2col.
// Replaces waveform_to_spectrogram (reference modular/source/inference_runner.py:157-174):
//   torch.stft framing (center, reflect pad 1024, 251 frames x 2048, hop 512)  -> stft_mel_kernel
//   periodic Hann window, rFFT, |X|^2                                           -> stft_mel_kernel
//   mel projection (1515 non-zero taps of the slaney/HTK filterbank)            -> stft_mel_kernel
//   10*log10(clamp(x,1e-10)); max(x, segment max - 80)                           -> stft_mel_kernel + db_clamp_stats_kernel
//   (x - mean) / (unbiased std + 1e-6)                                          -> db_clamp_stats_kernel + image_kernel
//   bilinear resize 128x251 -> 512x512 (anti-aliased taps), 3 identical channels -> image_kernel
#include <cuda_runtime.h>

#include <atomic>

#include "fft2048.cuh"
#include "frontend.h"

namespace sad {

namespace {

constexpr int kSeg = 128000;
constexpr int kFrames = 251;
constexpr int kMels = 128;
constexpr int kFramesPerCta = 32;
constexpr int kFrameGroups = (kFrames + kFramesPerCta - 1) / kFramesPerCta;   // 8

__device__ __forceinline__ int reflect_index(int j) {   // padded position -> source sample (pad 1024, reflect)
    int i = j - 1024;
    if (i < 0) i = -i;
    if (i >= kSeg) i = 2 * (kSeg - 1) - i;
    return i;
}

// Order-preserving float <-> uint mapping so a float max can use atomicMax on unsigned.
__device__ __forceinline__ unsigned f2ord(float f) {
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}

struct __align__(16) StftSmem {
    FftTwiddles tw;                 // 16 KB   per-pass twiddle tables, conflict-free [q][k] layout
    float re[kFftBuf];              // 9 KB
    float im[kFftBuf];              // 9 KB
    float pw[2][772];               // power spectra of the two frames, bins 0..768 (mel support is 2..768)
    float stage[kMels][kFramesPerCta + 1];   // dB staging, +1 pad
    float mel_w[1536];
    int mel_start[kMels];
    int mel_count[kMels];
    int mel_off[kMels];
    float red[8];
};

// grid (8 frame groups, B), 256 threads.  Writes UNCLAMPED dB to db[b][mel][frame] and the running per-segment
// maximum (ordered-uint encoding) to segmax[b].
__global__ void __launch_bounds__(kFftThreads) stft_mel_kernel(const float* __restrict__ pcm, const float* __restrict__ window,
                                                               const MelTable* __restrict__ mel, float* __restrict__ db,
                                                               unsigned* __restrict__ segmax) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    StftSmem& s = *reinterpret_cast<StftSmem*>(smem_raw);
    const int t = threadIdx.x;
    const int b = blockIdx.y;
    const int f0 = blockIdx.x * kFramesPerCta;
    const int nf = min(kFramesPerCta, kFrames - f0);
    const float* x = pcm + static_cast<size_t>(b) * kSeg;

    {
        cpx* flat = reinterpret_cast<cpx*>(&s.tw);
        for (int n = t; n < kFftTwiddleCount; n += kFftThreads) {
            float sn, cs;
            sincospif(static_cast<float>(fft_twiddle_angle(n)) * (1.0f / 1024.0f), &sn, &cs);
            flat[n] = {cs, -sn};
        }
    }
    for (int i = t; i < 1536; i += kFftThreads) s.mel_w[i] = i < mel->n_weights ? mel->w[i] : 0.f;
    if (t < kMels) {
        s.mel_start[t] = mel->start[t];
        s.mel_count[t] = mel->count[t];
        s.mel_off[t] = mel->off[t];
    }
    float win[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) win[q] = __ldg(window + t + 256 * q);
    __syncthreads();

    float local_max = -INFINITY;
    for (int fp = 0; fp < nf; fp += 2) {
        const int fa = f0 + fp;
        const bool has_b = (fp + 1) < nf;
        cpx in8[8];
        const bool interior = (fa >= 2) && (fa + 1 <= kFrames - 3);   // no reflection needed for either frame
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int n = t + 256 * q;
            const int ja = fa * 512 + n;
            float va, vb = 0.f;
            if (interior) {
                va = __ldg(x + ja - 1024);
                vb = __ldg(x + ja - 512);
            } else {
                va = __ldg(x + reflect_index(ja));
                if (has_b) vb = __ldg(x + reflect_index(ja + 512));
            }
            in8[q] = {va * win[q], vb * win[q]};
        }
        cpx v[8];
        fft_pass1(t, in8, s.re, s.im);
        __syncthreads();
        fft_pass2_load(t, s.re, s.im, s.tw, v);
        __syncthreads();
        fft_pass2_store(t, v, s.re, s.im);
        __syncthreads();
        fft_pass3_load(t, s.re, s.im, s.tw, v);
        __syncthreads();
        fft_pass3_store(t, v, s.re, s.im);
        __syncthreads();
        fft_pass4_load(t, s.re, s.im, s.tw, v);
        fft_pass4_store(t, v, s.re, s.im);   // same addresses this thread just read: no barrier needed
        __syncthreads();
        for (int k = t; k <= 768; k += kFftThreads) {
            float pa, pb;
            split_power(s.re, s.im, k, pa, pb);
            s.pw[0][k] = pa;
            s.pw[1][k] = pb;
        }
        __syncthreads();
        {
            const int which = t >> 7;      // 0: frame a, 1: frame b
            const int m = t & 127;
            if (which == 0 || has_b) {
                const float* p = s.pw[which] + s.mel_start[m];
                const float* w = s.mel_w + s.mel_off[m];
                float acc = 0.f;
                const int cnt = s.mel_count[m];
                for (int i = 0; i < cnt; ++i) acc = fmaf(p[i], w[i], acc);
                const float d = 10.0f * log10f(fmaxf(acc, 1e-10f));
                s.stage[m][fp + which] = d;
                local_max = fmaxf(local_max, d);
            }
        }
        // next iteration's pass-1 stores touch s.re/s.im only; s.pw is rewritten after 6 more barriers.
    }
    __syncthreads();
    // coalesced write-out: for each mel row, nf consecutive frames
    for (int i = t; i < kMels * kFramesPerCta; i += kFftThreads) {
        const int m = i / kFramesPerCta, f = i % kFramesPerCta;
        if (f < nf) db[(static_cast<size_t>(b) * kMels + m) * kFrames + f0 + f] = s.stage[m][f];
    }
    // block max -> one atomic per CTA
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local_max = fmaxf(local_max, __shfl_xor_sync(0xffffffffu, local_max, o));
    if ((t & 31) == 0) s.red[t >> 5] = local_max;
    __syncthreads();
    if (t == 0) {
        float m = s.red[0];
        for (int i = 1; i < 8; ++i) m = fmaxf(m, s.red[i]);
        atomicMax(segmax + b, f2ord(m));
    }
}

// One CTA per segment: clamp to (segment max - 80 dB), write the final log-mel dB (optional) and the
// mean / unbiased standard deviation over the 32128 cells.
__global__ void __launch_bounds__(512) db_clamp_stats_kernel(float* __restrict__ db, const unsigned* __restrict__ segmax,
                                                             float* __restrict__ out_db, float* __restrict__ mu_sigma,
                                                             float top_db) {
    constexpr int n = kMels * kFrames;
    const int b = blockIdx.x;
    float* src = db + static_cast<size_t>(b) * n;
    const float floor_db = ord2f(segmax[b]) - top_db;
    double s1 = 0.0, s2 = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float v = fmaxf(src[i], floor_db);
        src[i] = v;
        if (out_db) out_db[static_cast<size_t>(b) * n + i] = v;
        s1 += v;
        s2 += static_cast<double>(v) * v;
    }
    __shared__ double r1[16], r2[16];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if ((threadIdx.x & 31) == 0) {
        r1[threadIdx.x >> 5] = s1;
        r2[threadIdx.x >> 5] = s2;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0, q = 0;
        for (int i = 0; i < (blockDim.x >> 5); ++i) {
            a += r1[i];
            q += r2[i];
        }
        const double mean = a / n;
        double var = (q - a * mean) / (n - 1);
        if (var < 0) var = 0;
        mu_sigma[2 * b] = static_cast<float>(mean);
        mu_sigma[2 * b + 1] = static_cast<float>(sqrt(var));
    }
}

template <typename T>
__device__ __forceinline__ T to_out(float v);
template <>
__device__ __forceinline__ float to_out<float>(float v) { return v; }
template <>
__device__ __forceinline__ act_t to_out<act_t>(float v) { return act_from_float(v); }

// Standardise + separable 2-tap resize (horizontal first, taps accumulated with one fma each, as ATen does).
// grid (512 rows, B), 256 threads.  The two source rows of an output row are standardised ONCE into shared memory
// (502 IEEE divisions per CTA instead of 4 per output pixel), then every thread produces two output columns.
template <typename T>
__global__ void __launch_bounds__(256) image_kernel(const float* __restrict__ db, const float* __restrict__ mu_sigma,
                                                    const ResizeTable* __restrict__ rt, T* __restrict__ img) {
    __shared__ float rows[2][256];
    const int b = blockIdx.y, y = blockIdx.x;
    const float mu = mu_sigma[2 * b];
    const float den = mu_sigma[2 * b + 1] + 1e-6f;
    const float* src = db + static_cast<size_t>(b) * kMels * kFrames;
    const int y0 = rt->h_idx[y];
    const int y1 = min(y0 + 1, kMels - 1);
    const float wy0 = rt->h_w[2 * y], wy1 = rt->h_w[2 * y + 1];
    if (threadIdx.x < kFrames) {
        rows[0][threadIdx.x] = __fdiv_rn(src[y0 * kFrames + threadIdx.x] - mu, den);
        rows[1][threadIdx.x] = __fdiv_rn(src[y1 * kFrames + threadIdx.x] - mu, den);
    }
    __syncthreads();
    for (int x = threadIdx.x; x < 512; x += 256) {
        const int x0 = rt->w_idx[x];
        const int x1 = min(x0 + 1, kFrames - 1);
        const float wx0 = rt->w_w[2 * x], wx1 = rt->w_w[2 * x + 1];
        const float t0 = __fmaf_rn(rows[0][x1], wx1, __fmul_rn(rows[0][x0], wx0));
        const float t1 = __fmaf_rn(rows[1][x1], wx1, __fmul_rn(rows[1][x0], wx0));
        const float v = __fmaf_rn(t1, wy1, __fmul_rn(t0, wy0));
        img[(static_cast<size_t>(b) * 512 + y) * 512 + x] = to_out<T>(v);
    }
}

// Stem im2col for a generic 3-channel NCHW fp32 image: k = (ky*7+kx)*3 + c (147 real taps, zero to 192).
__global__ void __launch_bounds__(256) im2col_stem3_kernel(const float* __restrict__ x, act_t* __restrict__ A,
                                                           long long n_chunks) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n_chunks) return;
    const int chunk = static_cast<int>(i % 24);
    const long long pix = i / 24;
    const int ox = static_cast<int>(pix & 255);
    const int oy = static_cast<int>((pix >> 8) & 255);
    const long long b = pix >> 16;
    const float* src = x + b * 3 * 512 * 512;
    unsigned short v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int k = chunk * 8 + j;
        float val = 0.f;
        if (k < 147) {
            const int tap = k / 3, c = k % 3;
            const int iy = 2 * oy + tap / 7 - 3, ix = 2 * ox + tap % 7 - 3;
            if (iy >= 0 && iy < 512 && ix >= 0 && ix < 512) val = __ldg(src + (c * 512 + iy) * 512 + ix);
        }
        {
            const act_t a = act_from_float(val);
            v[j] = *reinterpret_cast<const unsigned short*>(&a);
        }
    }
    uint4 o;
    o.x = v[0] | (static_cast<unsigned>(v[1]) << 16);
    o.y = v[2] | (static_cast<unsigned>(v[3]) << 16);
    o.z = v[4] | (static_cast<unsigned>(v[5]) << 16);
    o.w = v[6] | (static_cast<unsigned>(v[7]) << 16);
    reinterpret_cast<uint4*>(A)[i] = o;
}

__device__ __forceinline__ uint4 bf16x8_max(uint4 a, uint4 b) {
    uint4 r;
    r.x = act_max2(a.x, b.x);
    r.y = act_max2(a.y, b.y);
    r.z = act_max2(a.z, b.z);
    r.w = act_max2(a.w, b.w);
    return r;
}

// maxpool 3x3 / stride 2 / pad 1 on NHWC bf16 [n,256,256,64] -> [n,128,128,64]; one thread per 8 channels.
__global__ void __launch_bounds__(256) maxpool_kernel(const act_t* __restrict__ in, act_t* __restrict__ out,
                                                      long long n_chunks) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n_chunks) return;
    const int c8 = static_cast<int>(i & 7);
    const long long pix = i >> 3;
    const int ox = static_cast<int>(pix & 127);
    const int oy = static_cast<int>((pix >> 7) & 127);
    const long long n = pix >> 14;
    const uint4* src = reinterpret_cast<const uint4*>(in) + n * 256 * 256 * 8;
    uint4 m;
    bool first = true;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
        const int iy = 2 * oy + dy;
        if (iy < 0 || iy >= 256) continue;
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
            const int ix = 2 * ox + dx;
            if (ix < 0 || ix >= 256) continue;
            const uint4 v = __ldg(src + (static_cast<long long>(iy) * 256 + ix) * 8 + c8);
            m = first ? v : bf16x8_max(m, v);
            first = false;
        }
    }
    reinterpret_cast<uint4*>(out)[i] = m;
}

// slice_waveform's silence gate: keep[w] = !(max |x| over the window < thr).  One CTA per window.
__global__ void __launch_bounds__(256) slice_gate_kernel(const float* __restrict__ wf, long long window, long long hop,
                                                         float thr, uint8_t* __restrict__ keep) {
    const float* x = wf + static_cast<long long>(blockIdx.x) * hop;
    // torch's max propagates NaN and `NaN < thr` is False (IR:186): a window holding a NaN is KEPT.  fmaxf drops NaNs,
    // so they are tracked separately.
    float m = 0.f;
    int nan = 0;
    for (long long i = threadIdx.x; i < window; i += blockDim.x) {
        const float v = fabsf(x[i]);
        nan |= (v != v);
        m = fmaxf(m, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    nan = __any_sync(0xffffffffu, nan);
    __shared__ float red[8];
    __shared__ int red_nan[8];
    if ((threadIdx.x & 31) == 0) {
        red[threadIdx.x >> 5] = m;
        red_nan[threadIdx.x >> 5] = nan;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; ++i) {
            m = fmaxf(m, red[i]);
            nan |= red_nan[i];
        }
        keep[blockIdx.x] = (!nan && m < thr) ? 0 : 1;
    }
}

__global__ void __launch_bounds__(256) gather_windows_kernel(const float* __restrict__ wf, const long long* __restrict__ starts,
                                                             long long window, float* __restrict__ dst) {
    // window index on gridDim.x (2^31-1 blocks): a long clip at a small hop has more than 65535 kept windows
    const long long w = blockIdx.x / 32, part = blockIdx.x % 32;
    const float* x = wf + starts[w];
    float* d = dst + w * window;
    for (long long i = part * blockDim.x + threadIdx.x; i < window; i += 32LL * blockDim.x) d[i] = x[i];
}

__global__ void fill_u32_kernel(unsigned* p, unsigned v, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

}  // namespace

size_t stft_smem_bytes() { return sizeof(StftSmem); }

cudaError_t frontend_logmel_launch(const float* pcm, int B, const float* window, const MelTable* mel, float* db_work,
                                   unsigned* segmax, float* out_db, float* mu_sigma, cudaStream_t stream, long long* launches) {
    {
        static std::atomic<bool> done[64];   // per device; two contexts may launch from two host threads
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        if (dev < 0 || dev >= 64 || !done[dev].load(std::memory_order_acquire)) {
            e = cudaFuncSetAttribute(stft_mel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(sizeof(StftSmem)));
            if (e != cudaSuccess) return e;
            if (dev >= 0 && dev < 64) done[dev].store(true, std::memory_order_release);
        }
    }
    fill_u32_kernel<<<(B + 255) / 256, 256, 0, stream>>>(segmax, 0u, B);   // 0 orders below every float
    stft_mel_kernel<<<dim3(kFrameGroups, B), kFftThreads, sizeof(StftSmem), stream>>>(pcm, window, mel, db_work, segmax);
    db_clamp_stats_kernel<<<B, 512, 0, stream>>>(db_work, segmax, out_db, mu_sigma, 80.0f);
    if (launches) *launches += 3;
    return cudaGetLastError();
}

cudaError_t image_launch_f32(const float* db, const float* mu_sigma, const ResizeTable* rt, float* img, int B,
                             cudaStream_t stream, long long* launches) {
    image_kernel<float><<<dim3(512, B), 256, 0, stream>>>(db, mu_sigma, rt, img);
    if (launches) *launches += 1;
    return cudaGetLastError();
}
cudaError_t image_launch_bf16(const float* db, const float* mu_sigma, const ResizeTable* rt, act_t* img, int B,
                              cudaStream_t stream, long long* launches) {
    image_kernel<act_t><<<dim3(512, B), 256, 0, stream>>>(db, mu_sigma, rt, img);
    if (launches) *launches += 1;
    return cudaGetLastError();
}
cudaError_t im2col_stem3_launch(const float* x, act_t* A, int B, cudaStream_t stream, long long* launches) {
    const long long n = static_cast<long long>(B) * 65536 * 24;
    im2col_stem3_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(x, A, n);
    if (launches) *launches += 1;
    return cudaGetLastError();
}
cudaError_t maxpool_launch(const act_t* in, act_t* out, long long n_img, cudaStream_t stream,
                           long long* launches) {
    const long long n = n_img * 128 * 128 * 8;
    maxpool_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(in, out, n);
    if (launches) *launches += 1;
    return cudaGetLastError();
}
cudaError_t slice_gate_launch(const float* wf, long long n_windows, long long window, long long hop, float thr,
                              uint8_t* keep, cudaStream_t stream, long long* launches) {
    if (n_windows <= 0) return cudaSuccess;
    slice_gate_kernel<<<static_cast<unsigned>(n_windows), 256, 0, stream>>>(wf, window, hop, thr, keep);
    if (launches) *launches += 1;
    return cudaGetLastError();
}
cudaError_t gather_windows_launch(const float* wf, const long long* starts, int n_kept, long long window, float* dst,
                                  cudaStream_t stream, long long* launches) {
    if (n_kept <= 0) return cudaSuccess;
    gather_windows_kernel<<<static_cast<unsigned>(n_kept) * 32u, 256, 0, stream>>>(wf, starts, window, dst);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

}  // namespace sad
