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
#include <cstdlib>

#include "fft2048.cuh"
#include "fft2048r16.cuh"
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

// ------------------------------------------------------------------------------------------------------------------
// logmel_kernel (round 2): the whole front end of a segment in ONE launch -- framing, Hann, 126 packed 2048-point FFTs
// (radix 16-16-8, fft2048r16.cuh), |X|^2, banded mel projection, 10*log10, the per-segment max for the top_db clamp,
// the clamp, mean / unbiased std, and the final [mel][frame] layout.  Replaces fill_u32 + stft_mel_kernel +
// db_clamp_stats_kernel (three launches, the unclamped dB written, re-read, re-written and read again by the image
// kernel, a global atomic max per CTA).
//   CTA = 896 threads = 7 groups of 128; a group transforms one frame PAIR at a time (126 pairs = 18 rounds x 7 groups,
//   no idle round) and synchronises on its own named barrier, so the groups hide each other's barrier and
//   shared-memory latency.  Persistent: CTA c handles segments c, c + gridDim.x, ...
//   Phase 1 writes unclamped dB FRAME-major to a scratch buffer (thread = mel band: coalesced 512-byte rows; the
//   segment's 128 KB stay in L2); phase 2 (same CTA, after a block-wide max) clamps, accumulates the statistics in
//   fp64 and transposes 32x32 tiles through shared memory into the [mel][frame] layout the reference produces.
// Build-time shape of the CTA, measured on B200 (bench.py --frontend-only, batch 4096, segments/s):
//   3 groups, constants in registers (168 regs), prefetch   783 k     6 groups (80 regs, 200 B spilled), prefetch   814 k
//   4 groups, tw3 + window in smem (128 regs), prefetch     865 k     6 groups (80 regs), no prefetch               868 k
//   5 groups, all constants in smem (96 regs), prefetch     883 k     7 groups (72 regs), no prefetch               914 k
// More resident warps beat holding constants / the next pair's samples in registers; at 7 groups (28 warps) the
// shared-memory pipe is the limit (~2 100 wavefronts per FFT against ~2 500 cycles).
#ifndef SAD_FE_GROUPS
#define SAD_FE_GROUPS 7
#endif
#ifndef SAD_FE_PREFETCH
#define SAD_FE_PREFETCH 0
#endif
constexpr int kFeGroups = SAD_FE_GROUPS;                // 128-thread FFT groups per CTA
constexpr int kFeThreads = kFft16Threads * kFeGroups;   // 896
constexpr int kPairs = (kFrames + 1) / 2;               // 126 frame pairs = 18 rounds x 7 groups (also even for 3 and 6)
constexpr bool kFeSmemConst = kFeGroups > 3;
constexpr bool kFeSmemTw2 = kFeGroups > 4;               // pass-2 twiddles in smem too (5+ groups: <= 102 registers)
constexpr bool kFePrefetch = SAD_FE_PREFETCH != 0;

struct __align__(16) FeSmem {
    cpx buf[kFeGroups][kFft16Slots];      // 17 KB per group; phase 2 reuses it as one [32][33]-float transposition tile per warp
    float pw[kFeGroups][2][772];          // power spectra of the group's two frames, bins 0..768 (769..771 stay zero)
    float mel_w[2816];                    // taps widened to 4-bin boundaries (MelTable::w4)
    cpx tw3[kFeSmemConst ? 14 : 1][kFft16Threads];   // [h*7 + q-1][t]: pass-3 twiddles when they do not fit in registers
    float win[kFeSmemConst ? 16 : 1][kFft16Threads]; // [q][t]: Hann window samples t + 128 q
    cpx tw2[kFeSmemTw2 ? 15 : 1][16];                // [q-1][k]: pass-2 twiddles (5+ groups)
    float red_max[kFeThreads / 32];
    double red_s1[kFeThreads / 32];
    double red_s2[kFeThreads / 32];
};
static_assert(sizeof(cpx) * kFeGroups * kFft16Slots >= (kFeThreads / 32) * 32 * 33 * sizeof(float), "tile aliasing");

__device__ __forceinline__ void group_bar(int g) { asm volatile("bar.sync %0, %1;\n" ::"r"(g + 1), "r"(kFft16Threads) : "memory"); }

__global__ void __launch_bounds__(kFeThreads, 1)
    logmel_kernel(const float* __restrict__ pcm, const float* __restrict__ window, const MelTable* __restrict__ mel,
                  float* __restrict__ scratch /*[B][251][128] unclamped dB*/, float* __restrict__ db /*[B][128][251]*/,
                  float* __restrict__ out_db /*optional copy*/, float* __restrict__ mu_sigma, int B, float top_db) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    FeSmem& s = *reinterpret_cast<FeSmem*>(smem_raw);
    const int tid = threadIdx.x;
    const int g = tid >> 7;
    const int t = tid & 127;
    const int warp = tid >> 5, lane = tid & 31;

    // per-thread constants of the whole kernel: twiddles (exact look-ups of exp(-2 pi i n / 2048)) and window samples
    cpx tw2[kFeSmemTw2 ? 1 : 15], tw3[kFeSmemConst ? 1 : 2][7];
    float win[kFeSmemConst ? 1 : 16];
#pragma unroll
    for (int q = 1; q < 16; ++q) {
        float sn, cs;
        sincospif(static_cast<float>(fft16_tw2_angle(t, q)) * (1.0f / 1024.0f), &sn, &cs);
        if constexpr (kFeSmemTw2) {
            if (tid < 16) s.tw2[q - 1][tid] = {cs, -sn};
        } else {
            tw2[q - 1] = {cs, -sn};
        }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int q = 1; q < 8; ++q) {
            float sn, cs;
            sincospif(static_cast<float>(fft16_tw3_angle(t, h, q)) * (1.0f / 1024.0f), &sn, &cs);
            if constexpr (kFeSmemConst) {
                if (g == 0) s.tw3[h * 7 + q - 1][t] = {cs, -sn};
            } else {
                tw3[h][q - 1] = {cs, -sn};
            }
        }
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        if constexpr (kFeSmemConst) {
            if (g == 0) s.win[q][t] = __ldg(window + t + 128 * q);
        } else {
            win[q] = __ldg(window + t + 128 * q);
        }
    }
    for (int i = tid; i < 2816; i += kFeThreads) s.mel_w[i] = mel->w4[i];
    for (int i = tid; i < kFeGroups * 2 * 772; i += kFeThreads) (&s.pw[0][0][0])[i] = 0.f;   // zero-weight taps may read bins 769..771
    __syncthreads();
    // mel band t for the pair's first frame, band 127-t for the second: the two tap counts sum to about the same for
    // every thread (2..36 taps per band, growing with the band index), so the four warps of a group finish together.
    // Taps and weights are read four at a time (16-byte loads on 4-bin boundaries; the widening taps have weight 0, which
    // leaves the sum bit-identical): scalar loads at per-lane offsets cost 3.3 wavefronts each (ncu, round 2).
    const int ma = t, mb = kMels - 1 - t;
    const int sa = __ldg(&mel->start4[ma]), ca = __ldg(&mel->count4[ma]), oa = __ldg(&mel->off4[ma]);
    const int sb = __ldg(&mel->start4[mb]), cb = __ldg(&mel->count4[mb]), ob = __ldg(&mel->off4[mb]);
    cpx* buf = s.buf[g];

    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        const float* x = pcm + static_cast<size_t>(b) * kSeg;
        float* scr = scratch + static_cast<size_t>(b) * kFrames * kMels;
        float local_max = -INFINITY;
        // raw samples of a frame pair: x[t + 128q] of frame 2p (ra) and of frame 2p+1 (rb); loaded one pair AHEAD (issued
        // after pass 3, consumed at the top of the next iteration) so that the L2 latency hides behind the power / mel stage
        float ra[16], rb[16];
        auto load_pair = [&](int p) {
            const int fa = 2 * p;
            const bool has_b = fa + 1 < kFrames;
            const bool interior = (fa >= 2) && (fa + 1 <= kFrames - 3);   // no reflection needed for either frame
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const int ja = fa * 512 + t + 128 * q;                   // position in the reflect-padded signal
                if (interior) {
                    ra[q] = __ldg(x + ja - 1024);
                    rb[q] = __ldg(x + ja - 512);
                } else {
                    ra[q] = __ldg(x + reflect_index(ja));
                    rb[q] = has_b ? __ldg(x + reflect_index(ja + 512)) : 0.f;
                }
            }
        };
        if (kFePrefetch) load_pair(g);
        for (int p = g; p < kPairs; p += kFeGroups) {
            const int fa = 2 * p;
            const bool has_b = fa + 1 < kFrames;
            if (!kFePrefetch) load_pair(p);
            cpx v[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const float w = kFeSmemConst ? s.win[kFeSmemConst ? q : 0][t] : win[kFeSmemConst ? 0 : q];
                v[q] = {ra[q] * w, rb[q] * w};
            }
            fft16_pass1(t, v, buf);
            group_bar(g);
            if constexpr (kFeSmemTw2) {
                fft16_pass2_tw(t, [&](int q) { return s.tw2[q - 1][t & 15]; }, buf);
            } else {
                fft16_pass2(t, tw2, buf);
            }
            group_bar(g);
            if constexpr (kFeSmemConst) {
                fft16_pass3_tw(t, [&](int h, int q) { return s.tw3[h * 7 + q - 1][t]; }, buf);
            } else {
                fft16_pass3(t, tw3, buf);
            }
            group_bar(g);
            if (kFePrefetch && p + kFeGroups < kPairs) load_pair(p + kFeGroups);
#pragma unroll
            for (int i = 0; i < 7; ++i) {
                const int k = t + 128 * i;
                if (k <= 768) {
                    float pa, pb;
                    fft16_split_power(buf, k, pa, pb);
                    s.pw[g][0][k] = pa;
                    s.pw[g][1][k] = pb;
                }
            }
            group_bar(g);
            {
                const float4* pp = reinterpret_cast<const float4*>(s.pw[g][0] + sa);
                const float4* ww = reinterpret_cast<const float4*>(s.mel_w + oa);
                float acc = 0.f;
                for (int i = 0; i < ca / 4; ++i) {
                    const float4 pv = pp[i], wv = ww[i];
                    acc = fmaf(pv.x, wv.x, acc);
                    acc = fmaf(pv.y, wv.y, acc);
                    acc = fmaf(pv.z, wv.z, acc);
                    acc = fmaf(pv.w, wv.w, acc);
                }
                const float d = 10.0f * log10f(fmaxf(acc, 1e-10f));
                scr[fa * kMels + ma] = d;
                local_max = fmaxf(local_max, d);
            }
            if (has_b) {
                const float4* pp = reinterpret_cast<const float4*>(s.pw[g][1] + sb);
                const float4* ww = reinterpret_cast<const float4*>(s.mel_w + ob);
                float acc = 0.f;
                for (int i = 0; i < cb / 4; ++i) {
                    const float4 pv = pp[i], wv = ww[i];
                    acc = fmaf(pv.x, wv.x, acc);
                    acc = fmaf(pv.y, wv.y, acc);
                    acc = fmaf(pv.z, wv.z, acc);
                    acc = fmaf(pv.w, wv.w, acc);
                }
                const float d = 10.0f * log10f(fmaxf(acc, 1e-10f));
                scr[(fa + 1) * kMels + mb] = d;
                local_max = fmaxf(local_max, d);
            }
            // the next pair's pass-1 stores only touch `buf` (all its readers passed the barrier above); pw is rewritten
            // after three more group barriers
        }
        // ---- phase 2: segment max -> clamp, statistics, [frame][mel] -> [mel][frame]
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) local_max = fmaxf(local_max, __shfl_xor_sync(0xffffffffu, local_max, o));
        if (lane == 0) s.red_max[warp] = local_max;
        __syncthreads();                                   // also orders this CTA's scratch writes before its reads below
        float seg_max = s.red_max[0];
#pragma unroll
        for (int i = 1; i < kFeThreads / 32; ++i) seg_max = fmaxf(seg_max, s.red_max[i]);
        const float floor_db = seg_max - top_db;
        float (*tile)[33] = reinterpret_cast<float (*)[33]>(reinterpret_cast<float*>(s.buf) + warp * 32 * 33);
        float* dst = db + static_cast<size_t>(b) * kMels * kFrames;
        float* dst2 = out_db ? out_db + static_cast<size_t>(b) * kMels * kFrames : nullptr;
        double s1 = 0.0, s2 = 0.0;
        // 8 frame blocks x 4 mel blocks of 32x32; warp w takes tiles w, w + 12, ...
        for (int tl = warp; tl < 32; tl += kFeThreads / 32) {
            const int f0 = (tl >> 2) * 32, m0 = (tl & 3) * 32;
#pragma unroll 8
            for (int r = 0; r < 32; ++r) {
                const int f = f0 + r;
                float vv = 0.f;
                if (f < kFrames) {
                    vv = fmaxf(scr[f * kMels + m0 + lane], floor_db);
                    s1 += vv;
                    s2 += static_cast<double>(vv) * vv;
                }
                tile[r][lane] = vv;
            }
            __syncwarp();
            const int f = f0 + lane;
#pragma unroll 8
            for (int r = 0; r < 32; ++r) {
                if (f < kFrames) {
                    const float vv = tile[lane][r];
                    dst[(m0 + r) * kFrames + f] = vv;
                    if (dst2) dst2[(m0 + r) * kFrames + f] = vv;
                }
            }
            __syncwarp();
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        if (lane == 0) {
            s.red_s1[warp] = s1;
            s.red_s2[warp] = s2;
        }
        __syncthreads();
        if (tid == 0) {
            constexpr int n = kMels * kFrames;
            double a = 0, q = 0;
            for (int i = 0; i < kFeThreads / 32; ++i) {
                a += s.red_s1[i];
                q += s.red_s2[i];
            }
            const double mean = a / n;
            double var = (q - a * mean) / (n - 1);
            if (var < 0) var = 0;
            mu_sigma[2 * b] = static_cast<float>(mean);
            mu_sigma[2 * b + 1] = static_cast<float>(sqrt(var));
        }
        __syncthreads();                                   // tiles alias the FFT buffers of the next segment
    }
}

// Standardise + separable 2-tap resize (horizontal first, taps accumulated with one fma each, as ATen does).
// grid (512 / kImgRows, B), 256 threads.  A CTA produces kImgRows = 8 consecutive output rows: they read at most 4 source
// rows (the vertical scale is 4), which are standardised ONCE into shared memory (<= 1004 IEEE divisions per CTA instead of
// 4 per output pixel; one CTA per output row spent most of its time on launch / drain: 65 536 CTAs per chunk).
// A thread owns the two output columns 2t and 2t+1 of all eight rows: their column taps stay in registers, the
// horizontal pass runs once per SOURCE row (4 x 2 values) instead of once per output row, and a row is written as one
// 4-byte (bf16 pair) or 8-byte (fp32 pair) store per thread.  Per pixel the arithmetic -- and so every bit -- is unchanged
// (round 2: 112 shared loads and 16 two-byte stores per thread before, 16 loads and 8 stores now).
constexpr int kImgRows = 8;
template <typename T>
__global__ void __launch_bounds__(256) image_kernel(const float* __restrict__ db, const float* __restrict__ mu_sigma,
                                                    const ResizeTable* __restrict__ rt, T* __restrict__ img) {
    __shared__ float rows[4][256];
    const int b = blockIdx.y, yb = blockIdx.x * kImgRows;
    const float mu = mu_sigma[2 * b];
    const float den = mu_sigma[2 * b + 1] + 1e-6f;
    const float* src = db + static_cast<size_t>(b) * kMels * kFrames;
    const int r_lo = rt->h_idx[yb];                                   // first source row any of the 8 output rows reads
    for (int i = threadIdx.x; i < 4 * 256; i += 256) {
        const int r = i >> 8, x = i & 255;
        const int sr = min(r_lo + r, kMels - 1);
        if (x < kFrames) rows[r][x] = __fdiv_rn(src[sr * kFrames + x] - mu, den);
    }
    // column taps of this thread's two pixels
    const int xa = 2 * threadIdx.x;
    int x0[2], x1[2];
    float wx0[2], wx1[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        x0[e] = rt->w_idx[xa + e];
        x1[e] = min(x0[e] + 1, kFrames - 1);
        wx0[e] = rt->w_w[2 * (xa + e)];
        wx1[e] = rt->w_w[2 * (xa + e) + 1];
    }
    __syncthreads();
    float h[4][2];                                                    // horizontal pass of the 4 source rows
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int e = 0; e < 2; ++e) h[r][e] = __fmaf_rn(rows[r][x1[e]], wx1[e], __fmul_rn(rows[r][x0[e]], wx0[e]));
#pragma unroll
    for (int ry = 0; ry < kImgRows; ++ry) {
        const int y = yb + ry;
        const int y0 = rt->h_idx[y];
        const int i0 = y0 - r_lo, i1 = min(y0 + 1, kMels - 1) - r_lo;  // 0 .. 3
        const float wy0 = rt->h_w[2 * y], wy1 = rt->h_w[2 * y + 1];
        float v[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            // h[i][e] with a run-time i: select instead of indexing (keeps h in registers)
            const float t0 = i0 == 0 ? h[0][e] : i0 == 1 ? h[1][e] : i0 == 2 ? h[2][e] : h[3][e];
            const float t1 = i1 == 0 ? h[0][e] : i1 == 1 ? h[1][e] : i1 == 2 ? h[2][e] : h[3][e];
            v[e] = __fmaf_rn(t1, wy1, __fmul_rn(t0, wy0));
        }
        T* o = img + (static_cast<size_t>(b) * 512 + y) * 512 + xa;
        if constexpr (sizeof(T) == 4) {
            *reinterpret_cast<float2*>(o) = make_float2(v[0], v[1]);
        } else {
            *reinterpret_cast<uint32_t*>(o) = act_pack(v[0], v[1]);
        }
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
                                   unsigned* segmax, float* out_db, float* mu_sigma, float* scratch, cudaStream_t stream,
                                   long long* launches) {
    static const bool v1 = [] {
        const char* e = getenv("SAD_FE_V1");     // A/B switch: the round-1 front end (three launches, radix 8-8-8-4)
        return e && atoi(e) != 0;
    }();
    static std::atomic<bool> done[64];   // per device; two contexts may launch from two host threads
    static std::atomic<int> sms[64];
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    if (!done[dev].load(std::memory_order_acquire)) {
        e = cudaFuncSetAttribute(stft_mel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(StftSmem)));
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(logmel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(FeSmem)));
        if (e != cudaSuccess) return e;
        int n = 0;
        e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
        sms[dev].store(n, std::memory_order_relaxed);
        done[dev].store(true, std::memory_order_release);
    }
    if (v1 || !scratch) {
        fill_u32_kernel<<<(B + 255) / 256, 256, 0, stream>>>(segmax, 0u, B);   // 0 orders below every float
        stft_mel_kernel<<<dim3(kFrameGroups, B), kFftThreads, sizeof(StftSmem), stream>>>(pcm, window, mel, db_work, segmax);
        db_clamp_stats_kernel<<<B, 512, 0, stream>>>(db_work, segmax, out_db, mu_sigma, 80.0f);
        if (launches) *launches += 3;
        return cudaGetLastError();
    }
    const int n_sm = sms[dev].load(std::memory_order_relaxed);
    logmel_kernel<<<B < n_sm ? B : n_sm, kFeThreads, sizeof(FeSmem), stream>>>(pcm, window, mel, scratch, db_work, out_db, mu_sigma,
                                                                             B, 80.0f);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t image_launch_f32(const float* db, const float* mu_sigma, const ResizeTable* rt, float* img, int B,
                             cudaStream_t stream, long long* launches) {
    image_kernel<float><<<dim3(512 / kImgRows, B), 256, 0, stream>>>(db, mu_sigma, rt, img);
    if (launches) *launches += 1;
    return cudaGetLastError();
}
cudaError_t image_launch_bf16(const float* db, const float* mu_sigma, const ResizeTable* rt, act_t* img, int B,
                              cudaStream_t stream, long long* launches) {
    image_kernel<act_t><<<dim3(512 / kImgRows, B), 256, 0, stream>>>(db, mu_sigma, rt, img);
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
