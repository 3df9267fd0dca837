// K0: ingest -- interleaved PCM (int16 or float32, any channel count, any integer sample rate) -> mono float32 at
// 32 kHz, zero-padded to at least one 4-s window.  HBM-bound: every input sample is read once, every output written once.
//
// Replaces preprocess_waveform of the reference after the container is parsed
// (modular/source/inference_runner.py:144-155): torchaudio.load's int16 -> float32 scaling (x / 32768), `wf.mean(dim=0)`,
// `torchaudio.transforms.Resample(sr, 32000)` with its defaults (sinc_interp_hann, lowpass_filter_width 6, rolloff 0.99)
// and the zero padding to one window.
//
// torchaudio evaluates y[m*new + p] = sum_k K[p][k] * xpad[m*orig + k] with a dense [new][2*width+orig] kernel (conv1d,
// stride orig).  All but ~12*orig/min(orig,new) of those taps sit where the Hann window argument is clamped and are
// below 1e-32: build_resample_taps (host) keeps the unclamped band per phase, the kernel below sums only that band.
//
//   Fast kernel (`phase` variant): a block owns blockDim * R consecutive outputs, blockDim a multiple of the number of
//   phases, so a thread keeps ONE phase: its <= T taps live in registers for all R outputs.  The mono-mixed input span
//   of the block is staged once in shared memory (coalesced reads of the interleaved PCM, each frame converted and mixed
//   once), so per output the SM does T shared-memory reads and T FMAs and HBM sees every byte once.
//   Fallback (`generic`): one output per thread, taps from global memory, for ratios with > 1024 phases or > 80 taps.
//   Measured (B200, 25 min of stereo int16): 1.8-2.2 TB/s for 44.1 / 48 / 96 / 16 kHz input, 5.8 TB/s when no resampling is
//   needed.  The FIR runs on CUDA cores at one shared-memory load per FMA: 18 LDS per output with 2-way conflicts caps the
//   kernel near 2.4 TB/s.  A variant with two adjacent outputs per thread on 8-byte window loads (4x fewer LDS wavefronts)
//   measured the SAME 1.8 TB/s at 96 registers / 2 blocks per SM -- stage -> sync -> compute leaves HBM idle while a block
//   computes -- so it was dropped; hoisting four staging loads ahead of their first use (more registers, fewer resident
//   blocks) was 20% SLOWER: occupancy, not per-thread latency, carries this kernel.  A persistent, double-buffered
//   (cp.async.bulk) version is the next step.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <vector>

#include "ingest.h"

namespace sad {

namespace {

constexpr int kBlock = 256;

template <typename T>
__device__ __forceinline__ float pcm_to_float(T v);
template <>
__device__ __forceinline__ float pcm_to_float<int16_t>(int16_t v) { return static_cast<float>(v) * (1.0f / 32768.0f); }
template <>
__device__ __forceinline__ float pcm_to_float<float>(float v) { return v; }

// mean over channels in channel order, as ATen's sum-then-divide on a [C, T] tensor
template <typename T>
__device__ __forceinline__ float mono_mix(const T* __restrict__ pcm, long long frame, int channels) {
    // one load per stereo frame when the stream's base allows it (a view that starts at an odd element is only
    // element aligned: the C ABI states no alignment requirement); x/2 == x*0.5 exactly
    if (channels == 2 && (reinterpret_cast<uintptr_t>(pcm) & (2 * sizeof(T) - 1)) == 0) {
        if constexpr (sizeof(T) == 2) {
            const short2 v = reinterpret_cast<const short2*>(pcm)[frame];
            return (static_cast<float>(v.x) * (1.0f / 32768.0f) + static_cast<float>(v.y) * (1.0f / 32768.0f)) * 0.5f;
        } else {
            const float2 v = reinterpret_cast<const float2*>(pcm)[frame];
            return (v.x + v.y) * 0.5f;
        }
    }
    const T* p = pcm + frame * channels;
    float s = pcm_to_float<T>(p[0]);
    for (int c = 1; c < channels; ++c) s += pcm_to_float<T>(p[c]);
    return channels == 1 ? s : s / static_cast<float>(channels);
}

// Four consecutive mono frames f .. f+3 (f a multiple of 4; frames outside [0, n_frames) read as zero).  Mono and stereo
// streams whose base is 16-byte aligned use one or two 8/16-byte loads: with 4 bytes per load a staging loop keeps too few
// bytes in flight to cover HBM latency (ncu: 45% of the stall samples on the first use of the loaded value).
template <typename T>
__device__ __forceinline__ float4 mono4(const T* __restrict__ pcm, long long f, int channels, long long n_frames, bool aligned) {
    if (aligned && f >= 0 && f + 3 < n_frames) {
        if constexpr (sizeof(T) == 2) {
            constexpr float s = 1.0f / 32768.0f;
            if (channels == 2) {
                const int4 v = *reinterpret_cast<const int4*>(pcm + f * 2);
                const int r[4] = {v.x, v.y, v.z, v.w};
                float o[4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    o[i] = (static_cast<float>(static_cast<short>(r[i] & 0xFFFF)) * s + static_cast<float>(static_cast<short>(r[i] >> 16)) * s) * 0.5f;
                return make_float4(o[0], o[1], o[2], o[3]);
            }
            if (channels == 1) {
                const int2 v = *reinterpret_cast<const int2*>(pcm + f);
                return make_float4(static_cast<float>(static_cast<short>(v.x & 0xFFFF)) * s, static_cast<float>(static_cast<short>(v.x >> 16)) * s,
                                   static_cast<float>(static_cast<short>(v.y & 0xFFFF)) * s, static_cast<float>(static_cast<short>(v.y >> 16)) * s);
            }
        } else {
            if (channels == 2) {
                const float4 a = *reinterpret_cast<const float4*>(pcm + f * 2), b = *reinterpret_cast<const float4*>(pcm + f * 2 + 4);
                return make_float4((a.x + a.y) * 0.5f, (a.z + a.w) * 0.5f, (b.x + b.y) * 0.5f, (b.z + b.w) * 0.5f);
            }
            if (channels == 1) return *reinterpret_cast<const float4*>(pcm + f);
        }
    }
    float o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i] = (f + i >= 0 && f + i < n_frames) ? mono_mix(pcm, f + i, channels) : 0.f;
    return make_float4(o[0], o[1], o[2], o[3]);
}

// sr_in == 32000: mix + pad only
template <typename T>
__global__ void __launch_bounds__(kBlock) ingest_copy_kernel(const T* __restrict__ pcm, long long n_frames, int channels,
                                                             float* __restrict__ out, long long out_len, bool aligned) {
    const long long j = (static_cast<long long>(blockIdx.x) * kBlock + threadIdx.x) * 4;
    if (j >= out_len) return;
    const float4 v = mono4(pcm, j, channels, n_frames, aligned);     // frames >= n_frames read as zero: the padding
    if (j + 3 < out_len && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
        *reinterpret_cast<float4*>(out + j) = v;
    } else {
        const float o[4] = {v.x, v.y, v.z, v.w};
        for (int i = 0; i < 4 && j + i < out_len; ++i) out[j + i] = o[i];
    }
}

template <typename T>
__global__ void __launch_bounds__(kBlock) ingest_resample_kernel(const T* __restrict__ pcm, long long n_frames, int channels,
                                                                 ResamplePlan plan, const int* __restrict__ tap_first,
                                                                 const float* __restrict__ tap_w, float* __restrict__ out,
                                                                 long long n_real, long long out_len) {
    extern __shared__ float span[];
    const long long j0 = static_cast<long long>(blockIdx.x) * kBlock;
    long long j_hi = j0 + kBlock - 1;
    if (j_hi >= n_real) j_hi = n_real - 1;
    if (j0 < n_real) {
        const long long m_lo = j0 / plan.new_f, m_hi = j_hi / plan.new_f;
        const long long a_lo = m_lo * plan.orig_f - plan.width;                       // first input frame of the span
        const int n_span = static_cast<int>((m_hi - m_lo) * plan.orig_f) + plan.taps_full;
        for (int i = threadIdx.x; i < n_span; i += kBlock) {
            const long long f = a_lo + i;
            span[i] = (f >= 0 && f < n_frames) ? mono_mix(pcm, f, channels) : 0.f;     // conv1d's zero padding
        }
        __syncthreads();
        const long long j = j0 + threadIdx.x;
        if (j < n_real) {
            const long long m = j / plan.new_f;
            const int p = static_cast<int>(j - m * plan.new_f);
            const float* w = tap_w + static_cast<size_t>(p) * plan.max_taps;
            const float* x = span + (m - m_lo) * plan.orig_f + tap_first[p];
            float acc = 0.f;
            for (int k = 0; k < plan.max_taps; ++k) acc = fmaf(w[k], x[k], acc);
            out[j] = acc;
        }
    }
    const long long j = j0 + threadIdx.x;
    if (j >= n_real && j < out_len) out[j] = 0.f;                                      // IR:150-154
}


// One phase per thread: outputs j = j0 + t + i * blockDim (blockDim % new_f == 0), i < rounds.
template <typename In, int T>
__global__ void __launch_bounds__(1024) ingest_resample_phase_kernel(const In* __restrict__ pcm, long long n_frames, int channels,
                                                                      ResamplePlan plan, const int* __restrict__ tap_first,
                                                                      const float* __restrict__ tap_w, float* __restrict__ out,
                                                                      long long n_real, long long out_len, int rounds, bool aligned) {
    extern __shared__ float4 span4[];
    float* span = reinterpret_cast<float*>(span4);
    const int nt = blockDim.x;
    const int q = nt / plan.new_f;                                   // frames (of new_f outputs) per round
    const long long per_block = static_cast<long long>(nt) * rounds;
    const long long j0 = static_cast<long long>(blockIdx.x) * per_block;   // a multiple of new_f: phase 0 of frame m_base
    const long long m_base = static_cast<long long>(blockIdx.x) * q * rounds;
    long long j_hi = j0 + per_block - 1;
    if (j_hi >= n_real) j_hi = n_real - 1;
    if (j0 < n_real) {
        // first / last input frame any output of the block touches (positions are non-decreasing in j)
        const long long m1 = j_hi / plan.new_f;
        const int first0 = tap_first[0];
        const long long a_need = m_base * plan.orig_f + first0 - plan.width;
        const long long a_lo = a_need & ~3LL;                        // staged in groups of 4 frames (vector loads)
        const int lead = static_cast<int>(a_need - a_lo);
        const int n_span = lead + static_cast<int>((m1 - m_base) * plan.orig_f) + tap_first[j_hi - m1 * plan.new_f] - first0 + T;
        for (int i = threadIdx.x; i < (n_span + 3) / 4; i += nt)                      // (64-register cap here: no load hoisting)
            span4[i] = mono4(pcm, a_lo + 4LL * i, channels, n_frames, aligned);
        const int tm = threadIdx.x / plan.new_f;                     // 32-bit: frame of this thread inside a round
        const int p = threadIdx.x - tm * plan.new_f;
        float w[T];
#pragma unroll
        for (int k = 0; k < T; ++k) w[k] = k < plan.max_taps ? tap_w[static_cast<size_t>(p) * plan.max_taps + k] : 0.f;
        const int off0 = lead + tm * plan.orig_f + tap_first[p] - first0;   // offset of this thread's band in round 0
        const int rel_hi = static_cast<int>(j_hi - j0), t = static_cast<int>(threadIdx.x);
        const int n_mine = rel_hi >= t ? (rel_hi - t) / nt + 1 : 0;   // outputs of this thread that are real samples
        __syncthreads();
        for (int i = 0; i < rounds; ++i) {
            if (i >= n_mine) break;
            const float* x = span + off0 + i * q * plan.orig_f;
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < T; ++k) acc = fmaf(w[k], x[k], acc);
            out[j0 + threadIdx.x + static_cast<long long>(i) * nt] = acc;
        }
    }
    if (j0 + per_block > n_real)
        for (int i = 0; i < rounds; ++i) {                                             // IR:150-154
            const long long j = j0 + threadIdx.x + static_cast<long long>(i) * nt;
            if (j >= n_real && j < out_len) out[j] = 0.f;
        }
}

// ---- `quad` variant: four adjacent outputs per thread, persistent blocks, the stream staged by cp.async ------------------
//
// A thread's four outputs j .. j+3 read bands that start s_0 <= s_1 <= s_2 <= s_3 <= s_0 + quad_shift_max frames into the
// staged span, so ONE window of TE floats -- read as TE/4 aligned 16-byte loads -- feeds all four: 7 LDS.128 for four
// outputs in place of 4 x 18 LDS.32.  Each output's taps are kept in registers shifted to its place in that window
// (W[g][sh_g + k] = w_g[k], zero elsewhere: the products with zero weights are exact, the sum order of the other
// kernels is kept, so the results are the same to the bit).
//
// The raw interleaved PCM of the next kDepth items is in flight (cp.async, 16 bytes per request, zero-filled outside the
// stream) into a ring of shared buffers while the block filters the current item, so HBM stays busy with ONE resident
// block per SM (the tap registers allow no more) -- with the loads staged through registers instead, the first use of a
// loaded value stalled the warp in front of the filter loop and the kernel ran at 0.6x of the one-phase kernel.  Per
// item: wait for its raw chunks, convert + mix them into the float span (index 0 = the first frame the item needs, so a
// thread's window alignment never changes), refill the freed raw slot, filter.  Mono or stereo streams whose base is
// 16-byte aligned; everything else keeps the one-phase kernel.
constexpr int kQuadDepth = 3;

__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// one 16-byte chunk of the stream -> its mono frames (the arithmetic of mono4); returns how many
template <typename In>
__device__ __forceinline__ int chunk_to_mono(const int4 v, int channels, float (&o)[8]) {
    const int r[4] = {v.x, v.y, v.z, v.w};
    if constexpr (sizeof(In) == 2) {
        constexpr float s = 1.0f / 32768.0f;
        if (channels == 2) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                o[i] = (static_cast<float>(static_cast<short>(r[i] & 0xFFFF)) * s + static_cast<float>(static_cast<short>(r[i] >> 16)) * s) * 0.5f;
            return 4;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            o[2 * i] = static_cast<float>(static_cast<short>(r[i] & 0xFFFF)) * s;
            o[2 * i + 1] = static_cast<float>(static_cast<short>(r[i] >> 16)) * s;
        }
        return 8;
    } else {
        if (channels == 2) {
            o[0] = (__int_as_float(r[0]) + __int_as_float(r[1])) * 0.5f;
            o[1] = (__int_as_float(r[2]) + __int_as_float(r[3])) * 0.5f;
            return 2;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] = __int_as_float(r[i]);
        return 4;
    }
}

template <typename In, int TE>
__global__ void __launch_bounds__(384, 1) ingest_resample_quad_kernel(const In* __restrict__ pcm, long long n_frames, int channels,
                                                                       ResamplePlan plan, const int* __restrict__ tap_first,
                                                                       const float* __restrict__ tap_w, float* __restrict__ out,
                                                                       long long n_real, long long out_len, int rounds,
                                                                       int span_floats, int n_chunks, long long n_items, bool out16) {
    extern __shared__ float4 smem4[];
    float* const span = reinterpret_cast<float*>(smem4);             // [span_floats], a multiple of 4
    const int4* const raw = reinterpret_cast<const int4*>(smem4) + span_floats / 4;   // [kQuadDepth][n_chunks]
    const uint32_t raw_addr = static_cast<uint32_t>(__cvta_generic_to_shared(raw));
    const int nt = blockDim.x, t = threadIdx.x;
    const int bpf = static_cast<int>(sizeof(In)) * channels;        // bytes per frame: 2, 4 or 8
    const int fpc = 16 / bpf;                                       // frames per 16-byte chunk
    const int frames_round = 4 * nt / plan.new_f;                   // 4 * nt is a multiple of new_f
    const int round_stride = frames_round * plan.orig_f;            // floats; a multiple of 4 when rounds > 1
    const long long per_item = 4LL * nt * rounds;
    const long long frames_item = static_cast<long long>(frames_round) * rounds;
    if (static_cast<long long>(blockIdx.x) >= n_items) return;
    const int items_mine = static_cast<int>((n_items - blockIdx.x + gridDim.x - 1) / gridDim.x);
    const char* const bytes = reinterpret_cast<const char*>(pcm);

    auto first_frame = [&](long long it) { return it * frames_item * plan.orig_f + plan.first0 - plan.width; };
    auto issue = [&](int n, int slot) {                              // raw chunks of this block's n-th item -> ring slot
        const long long a_lo = first_frame(blockIdx.x + static_cast<long long>(n) * gridDim.x) & ~static_cast<long long>(fpc - 1);
        for (int i = t; i < n_chunks; i += nt) {
            const long long f = a_lo + static_cast<long long>(fpc) * i;
            long long nb = f < 0 ? 0 : (n_frames - f) * bpf;        // bytes of the stream from frame f on
            nb = nb < 0 ? 0 : nb > 16 ? 16 : nb;
            cp_async16_zfill(raw_addr + static_cast<uint32_t>(slot * n_chunks + i) * 16u, nb > 0 ? bytes + f * bpf : bytes,
                             static_cast<int>(nb));
        }
    };
#pragma unroll
    for (int d = 0; d < kQuadDepth; ++d) {
        if (d < items_mine) issue(d, d);
        cp_async_commit();
    }

    // this thread's window and the four shifted tap sets
    int pos[4], ph[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        const int m = (4 * t + g) / plan.new_f;
        ph[g] = (4 * t + g) - m * plan.new_f;
        pos[g] = m * plan.orig_f + tap_first[ph[g]] - plan.first0;
    }
    const int base = pos[0] & ~3;
    float W[4][TE];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        const int sh = pos[g] - base;
        const float* w = tap_w + static_cast<size_t>(ph[g]) * plan.max_taps;
#pragma unroll
        for (int i = 0; i < TE; ++i) {
            const int k = i - sh;
            W[g][i] = (k >= 0 && k < plan.max_taps) ? w[k] : 0.f;
        }
    }

    int slot = 0;
    for (int n = 0; n < items_mine; ++n) {
        const long long item = blockIdx.x + static_cast<long long>(n) * gridDim.x;
        cp_async_wait<kQuadDepth - 1>();                            // this thread's chunks of item n have landed
        __syncthreads();                                            // ... and everyone's; the previous filter pass is over
        {
            const int skip = static_cast<int>(first_frame(item) & static_cast<long long>(fpc - 1));
            const int4* src = raw + slot * n_chunks;
            for (int i = t; i < n_chunks; i += nt) {
                float o[8];
                const int cnt = chunk_to_mono<In>(src[i], channels, o);
                const int s0 = fpc * i - skip;
#pragma unroll
                for (int e = 0; e < 8; ++e)
                    if (e < cnt && s0 + e >= 0 && s0 + e < span_floats) span[s0 + e] = o[e];
            }
        }
        __syncthreads();
        if (n + kQuadDepth < items_mine) issue(n + kQuadDepth, slot);
        cp_async_commit();
        slot = slot + 1 == kQuadDepth ? 0 : slot + 1;

        const long long j_item = item * per_item + 4LL * t;
        for (int i = 0; i < rounds; ++i) {
            const float4* x = reinterpret_cast<const float4*>(span + base + i * round_stride);
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int c = 0; c < TE / 4; ++c) {
                const float4 v = x[c];
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    acc[g] = fmaf(W[g][4 * c + 0], v.x, acc[g]);
                    acc[g] = fmaf(W[g][4 * c + 1], v.y, acc[g]);
                    acc[g] = fmaf(W[g][4 * c + 2], v.z, acc[g]);
                    acc[g] = fmaf(W[g][4 * c + 3], v.w, acc[g]);
                }
            }
            const long long j = j_item + 4LL * nt * i;
#pragma unroll
            for (int g = 0; g < 4; ++g)
                if (j + g >= n_real) acc[g] = 0.f;                                     // IR:150-154
            if (out16 && j + 3 < out_len) {
                *reinterpret_cast<float4*>(out + j) = make_float4(acc[0], acc[1], acc[2], acc[3]);
            } else {
#pragma unroll
                for (int g = 0; g < 4; ++g)
                    if (j + g < out_len) out[j + g] = acc[g];
            }
        }
    }
    cp_async_wait<0>();
}

template <typename In, int TE>
cudaError_t launch_quad(const In* pcm, long long n_frames, int channels, const ResamplePlan& plan, const int* tap_first,
                        const float* tap_w, float* out, long long n_real, long long out_len, int threads, int rounds,
                        int span_floats, int n_chunks, cudaStream_t stream) {
    const size_t smem = static_cast<size_t>(span_floats) * sizeof(float) + static_cast<size_t>(kQuadDepth) * n_chunks * 16;
    const long long per_item = 4LL * threads * rounds;
    const long long n_items = (out_len + per_item - 1) / per_item;
    int dev = 0, sms = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(ingest_resample_quad_kernel<In, TE>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    const unsigned grid = static_cast<unsigned>(n_items < sms ? n_items : sms);      // one block per SM (168 registers x 320 threads)
    const bool out16 = (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    ingest_resample_quad_kernel<In, TE><<<grid, threads, smem, stream>>>(pcm, n_frames, channels, plan, tap_first, tap_w, out, n_real,
                                                                         out_len, rounds, span_floats, n_chunks, n_items, out16);
    return cudaGetLastError();
}

// The quad kernel applies when the window of four adjacent outputs (alignment slack 3 + quad_shift_max + taps) fits 28
// floats (the 32-float instantiation spills), a block of <= 384 threads holds whole periods of the phase pattern and the
// stream is mono or stereo on a 16-byte aligned base; SAD_INGEST_QUAD=0 turns it off.
template <typename In>
bool try_quad(const In* pcm, long long n_frames, int channels, const ResamplePlan& plan, const int* tap_first, const float* tap_w,
              float* out, long long n_real, long long out_len, cudaStream_t stream, cudaError_t* err) {
    static const bool enabled = [] { const char* v = getenv("SAD_INGEST_QUAD"); return !(v && v[0] == '0'); }();
    if (!enabled || plan.quad_shift_max < 0 || channels > 2 || n_frames < 1 || (reinterpret_cast<uintptr_t>(pcm) & 15) != 0) return false;
    const int need = 3 + plan.quad_shift_max + plan.max_taps;
    if (need > 28) return false;
    const int TE = need <= 20 ? 20 : need <= 24 ? 24 : 28;
    int period = plan.new_f;                                         // threads per period of the phase pattern: new_f / gcd(new_f, 4)
    for (int g = 0; g < 2 && period % 2 == 0; ++g) period /= 2;
    int threads = 0;
    for (int lo : {256, 128}) {
        threads = (lo + period - 1) / period * period;
        if (threads <= 384) break;
        threads = 0;
    }
    if (!threads) return false;
    const int fpc = 16 / (static_cast<int>(sizeof(In)) * channels);
    const int frames_round = 4 * threads / plan.new_f;
    const int max_rounds = (frames_round * plan.orig_f) % 4 == 0 ? 8 : 1;
    int rounds = max_rounds, span_floats = 0, n_chunks = 0;
    auto size_for = [&](int r) {
        const long long need_floats = (static_cast<long long>(frames_round) * r - 1) * plan.orig_f + plan.first_spread + TE;
        span_floats = static_cast<int>((need_floats + 3) / 4 * 4);
        n_chunks = static_cast<int>((need_floats + fpc - 1 + fpc - 1) / fpc);       // the first chunk may start fpc - 1 frames early
        return static_cast<size_t>(span_floats) * 4 + static_cast<size_t>(kQuadDepth) * n_chunks * 16;
    };
    while (rounds >= 1 && size_for(rounds) > 120 * 1024) rounds >>= 1;
    // short streams: smaller items, so that every SM gets a few
    while (rounds > 1 && (out_len + 4LL * threads * rounds - 1) / (4LL * threads * rounds) < 4 * 148) rounds >>= 1;
    if (rounds >= 1) size_for(rounds);
    if (rounds < 1) return false;
    switch (TE) {
        case 20: *err = launch_quad<In, 20>(pcm, n_frames, channels, plan, tap_first, tap_w, out, n_real, out_len, threads, rounds, span_floats, n_chunks, stream); break;
        case 24: *err = launch_quad<In, 24>(pcm, n_frames, channels, plan, tap_first, tap_w, out, n_real, out_len, threads, rounds, span_floats, n_chunks, stream); break;
        default: *err = launch_quad<In, 28>(pcm, n_frames, channels, plan, tap_first, tap_w, out, n_real, out_len, threads, rounds, span_floats, n_chunks, stream); break;
    }
    return true;
}

template <typename In, int T>
cudaError_t launch_phase(const In* pcm, long long n_frames, int channels, const ResamplePlan& plan, const int* tap_first,
                         const float* tap_w, float* out, long long n_real, long long out_len, int threads, int rounds,
                         size_t smem, cudaStream_t stream) {
    const long long per_block = static_cast<long long>(threads) * rounds;
    const unsigned grid = static_cast<unsigned>((out_len + per_block - 1) / per_block);
    const bool aligned = (reinterpret_cast<uintptr_t>(pcm) & 15) == 0;
    ingest_resample_phase_kernel<In, T><<<grid, threads, smem, stream>>>(pcm, n_frames, channels, plan, tap_first, tap_w, out,
                                                                         n_real, out_len, rounds, aligned);
    return cudaGetLastError();
}

template <typename In>
bool try_phase(const In* pcm, long long n_frames, int channels, const ResamplePlan& plan, const int* tap_first,
               const float* tap_w, float* out, long long n_real, long long out_len, cudaStream_t stream, cudaError_t* err) {
    if (plan.new_f > 1024 || plan.max_taps > 80) return false;
    const int T = plan.max_taps <= 14 ? 14 : plan.max_taps <= 18 ? 18 : plan.max_taps <= 20 ? 20 : plan.max_taps <= 24 ? 24 : plan.max_taps <= 40 ? 40 : 80;
    int threads = plan.new_f >= 256 ? plan.new_f : plan.new_f * (256 / plan.new_f);
    // inputs touched by threads*rounds consecutive outputs: ceil(outputs * orig / new) + T (+ slack for the band start)
    int rounds = 8;
    size_t smem = 0;
    for (; rounds >= 1; rounds >>= 1) {
        smem = (static_cast<size_t>(threads) * rounds * plan.orig_f / plan.new_f + plan.orig_f + 2 * T + 16) * sizeof(float);
        if (smem <= 46 * 1024) break;
    }
    if (rounds < 1) return false;
    switch (T) {
        case 14: *err = launch_phase<In, 14>(pcm, n_frames, channels, plan, tap_first, tap_w, out, n_real, out_len, threads, rounds, smem, stream); break;
        case 18: *err = launch_phase<In, 18>(pcm, n_frames, channels, plan, tap_first, tap_w, out, n_real, out_len, threads, rounds, smem, stream); break;
        case 20: *err = launch_phase<In, 20>(pcm, n_frames, channels, plan, tap_first, tap_w, out, n_real, out_len, threads, rounds, smem, stream); break;
        case 24: *err = launch_phase<In, 24>(pcm, n_frames, channels, plan, tap_first, tap_w, out, n_real, out_len, threads, rounds, smem, stream); break;
        case 40: *err = launch_phase<In, 40>(pcm, n_frames, channels, plan, tap_first, tap_w, out, n_real, out_len, threads, rounds, smem, stream); break;
        default: *err = launch_phase<In, 80>(pcm, n_frames, channels, plan, tap_first, tap_w, out, n_real, out_len, threads, rounds, smem, stream); break;
    }
    return true;
}


}  // namespace

size_t ingest_smem_bytes(const ResamplePlan& plan) {
    const long long m_span = (kBlock - 1) / plan.new_f + 1;             // m_hi - m_lo <= this
    return static_cast<size_t>(m_span * plan.orig_f + plan.taps_full) * sizeof(float);
}

cudaError_t ingest_launch(const void* pcm, int sample_format, long long n_frames, int channels, const ResamplePlan* plan,
                          const IngestTables& tb, float* out, long long n_real, long long out_len, cudaStream_t stream,
                          long long* launches) {
    const int* tap_first = tb.tap_first;
    const float* tap_w = tb.tap_w;
    if (out_len <= 0) return cudaSuccess;
    const unsigned grid = static_cast<unsigned>((out_len + kBlock - 1) / kBlock);
    if (!plan) {
        const unsigned grid4 = static_cast<unsigned>((out_len + 4 * kBlock - 1) / (4 * kBlock));
        const bool aligned = (reinterpret_cast<uintptr_t>(pcm) & 15) == 0;
        if (sample_format == 0)
            ingest_copy_kernel<int16_t><<<grid4, kBlock, 0, stream>>>(static_cast<const int16_t*>(pcm), n_frames, channels, out, out_len, aligned);
        else
            ingest_copy_kernel<float><<<grid4, kBlock, 0, stream>>>(static_cast<const float*>(pcm), n_frames, channels, out, out_len, aligned);
    } else {
        cudaError_t e = cudaSuccess;
        const bool quad = sample_format == 0
            ? try_quad(static_cast<const int16_t*>(pcm), n_frames, channels, *plan, tap_first, tap_w, out, n_real, out_len, stream, &e)
            : try_quad(static_cast<const float*>(pcm), n_frames, channels, *plan, tap_first, tap_w, out, n_real, out_len, stream, &e);
        if (quad) {
            if (launches) *launches += 1;
            return e;
        }
        const bool fast = sample_format == 0
            ? try_phase(static_cast<const int16_t*>(pcm), n_frames, channels, *plan, tap_first, tap_w, out, n_real, out_len, stream, &e)
            : try_phase(static_cast<const float*>(pcm), n_frames, channels, *plan, tap_first, tap_w, out, n_real, out_len, stream, &e);
        if (fast) {
            if (launches) *launches += 1;
            return e;
        }
        const size_t smem = ingest_smem_bytes(*plan);
        if (sample_format == 0)
            ingest_resample_kernel<int16_t><<<grid, kBlock, smem, stream>>>(static_cast<const int16_t*>(pcm), n_frames, channels,
                                                                            *plan, tap_first, tap_w, out, n_real, out_len);
        else
            ingest_resample_kernel<float><<<grid, kBlock, smem, stream>>>(static_cast<const float*>(pcm), n_frames, channels, *plan,
                                                                          tap_first, tap_w, out, n_real, out_len);
    }
    if (launches) *launches += 1;
    return cudaGetLastError();
}

}  // namespace sad
