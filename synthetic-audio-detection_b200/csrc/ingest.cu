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
//   `pair` kernel (every rate up to 96 kHz; mono or stereo on a 16-byte aligned base): persistent blocks, the
//   raw PCM of the next item staged by cp.async while the current one is filtered, two adjacent outputs per thread
//   on one window of 16-byte shared loads with the taps in registers.  Measured (B200, 25 min of stereo int16,
//   tools/ingest_bench.py): 3.8 TB/s at 44.1 kHz (58 % of the measured HBM copy bandwidth), 3.1-3.9 TB/s at 8 / 22.05 /
//   24 kHz, 2.2-2.4 TB/s at 88.2 kHz (one output per thread there); the few-phase form of the kernel (taps as kernel
//   parameters, ingest_taps.h UniformTaps) runs 48 kHz at 4.5 (69 %), 16 kHz at 5.1 (78 %), 96 kHz at 4.0 TB/s (61 %).  ncu on the way there: the first cut was instruction-issue bound (73 % issue slots, 131 instructions per
//   output against 24 useful FMAs: a conversion pass with a division and two guarded stores per frame, 64-bit index
//   arithmetic per round, the shared base address rebuilt from S2R at every use); the version below executes ~45.
//   `phase` kernel (other channel counts, unaligned streams, 192 kHz): a block owns blockDim * R consecutive outputs,
//   blockDim a multiple of the number of phases, so a thread keeps ONE phase: its <= T taps live in registers for all R
//   outputs.  The mono-mixed input span of the block is staged once in shared memory (coalesced reads of the interleaved
//   PCM, each frame converted and mixed once): T shared loads and T FMAs per output, stage -> sync -> compute.
//   1.8-2.1 TB/s; it was what every rate ran on before (18 LDS per output with 2-way conflicts, HBM idle while a block
//   computes).
//   `generic` fallback: one output per thread, taps from global memory, for ratios with > 1024 phases or > 80 taps.
//   No resampling (32 kHz input): mix + pad only, 5.9 TB/s (90 %).
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

// ---- `pair` variant: two adjacent outputs per thread, persistent blocks, the stream staged by cp.async -------------------
//
// A thread's outputs j and j+1 read bands that start s_0 <= s_1 <= s_0 + pair_shift_max frames into the staged span, so
// ONE window of TE floats -- read as TE/4 aligned 16-byte loads -- feeds both: 6 LDS.128 for two outputs in place of
// 2 x 18 LDS.32, and a quarter warp's windows lie within 128 bytes of each other (no bank conflicts; with four outputs
// per thread they span 176 and every load costs two wavefronts).  Each output's taps are kept in registers shifted to
// its place in that window (W[g][sh_g + k] = w_g[k], zero elsewhere: the products with zero weights are exact and the
// sum order of the other kernels is kept, so the results are the same to the bit).
//
// The raw interleaved PCM of the next item(s) (kPairDepth slots) is in flight (cp.async, 16 bytes per request, zero-filled outside
// the stream) into a ring of shared buffers while the block filters the current item -- with the loads staged through
// registers instead, the first use of a loaded value stalled the warp in front of the filter loop.  Per item: wait for
// its raw chunks, convert + mix them into the float span, refill the freed raw slot, filter.  The span holds one
// sub-span per round (the outputs 2 * blockDim apart), each starting on a 16-byte boundary with the few frames it shares
// with the next one written twice, so that a thread's window alignment is the same in every round and item whatever
// the rate ratio.  Two blocks per SM (<= 102 registers) overlap one block's conversion with the other's filter pass.
// Mono or stereo streams whose base is 16-byte aligned; everything else keeps the one-phase kernel.

__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void st_shared_f32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;\n" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ float4 ld_shared_f32x4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// One frame of the staged raw PCM (shared address) as a mono float.  Every operation is exact for int16 (a sum of two
// 16-bit values scaled by a power of two), so this equals mono_mix to the bit; the float mix is the same (a + b) * 0.5f.
template <typename In, int CH>
__device__ __forceinline__ float raw_frame_to_mono(uint32_t addr) {
    if constexpr (sizeof(In) == 4 && CH == 2) {
        float a, b;
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];\n" : "=f"(a), "=f"(b) : "r"(addr) : "memory");
        return (a + b) * 0.5f;
    } else if constexpr (sizeof(In) == 4) {
        float a;
        asm volatile("ld.shared.f32 %0, [%1];\n" : "=f"(a) : "r"(addr) : "memory");
        return a;
    } else if constexpr (CH == 2) {
        int v;
        asm volatile("ld.shared.b32 %0, [%1];\n" : "=r"(v) : "r"(addr) : "memory");
        return static_cast<float>(__dp2a_lo(v, 0x0101, 0)) * (0.5f / 32768.0f);                 // lo + hi in one instruction
    } else {
        int v;
        asm volatile("ld.shared.s16 %0, [%1];\n" : "=r"(v) : "r"(addr) : "memory");
        return static_cast<float>(v) * (1.0f / 32768.0f);
    }
}

// Two consecutive frames (the first one at an even index: 2 * frame-size aligned) -> two mono floats, same arithmetic.
template <typename In, int CH>
__device__ __forceinline__ float2 raw_two_frames_to_mono(uint32_t addr) {
    if constexpr (sizeof(In) == 4 && CH == 2) {
        float a, b, c, d;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n" : "=f"(a), "=f"(b), "=f"(c), "=f"(d) : "r"(addr) : "memory");
        return make_float2((a + b) * 0.5f, (c + d) * 0.5f);
    } else if constexpr (sizeof(In) == 4) {
        float a, b;
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];\n" : "=f"(a), "=f"(b) : "r"(addr) : "memory");
        return make_float2(a, b);
    } else if constexpr (CH == 2) {
        int v, w;
        asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];\n" : "=r"(v), "=r"(w) : "r"(addr) : "memory");
        return make_float2(static_cast<float>(__dp2a_lo(v, 0x0101, 0)) * (0.5f / 32768.0f),
                           static_cast<float>(__dp2a_lo(w, 0x0101, 0)) * (0.5f / 32768.0f));
    } else {
        int v;
        asm volatile("ld.shared.b32 %0, [%1];\n" : "=r"(v) : "r"(addr) : "memory");
        return make_float2(static_cast<float>(static_cast<short>(v & 0xFFFF)) * (1.0f / 32768.0f),
                           static_cast<float>(v >> 16) * (1.0f / 32768.0f));
    }
}
__device__ __forceinline__ void st_shared_f32x2(uint32_t addr, float2 v) {
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};\n" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory");
}

// UORIG > 0 selects the few-phase form (ingest_taps.h UniformTaps): UNEW phases, UTAPS shifted taps per phase read from
// the kernel parameter `uni` as constant-bank operands, G outputs per thread on a window that starts at t * (G / UNEW) *
// UORIG -- no tap registers, four blocks per SM.
template <typename In, int CH, int TE, int G, int MINB = 2, int UORIG = 0, int UNEW = 1, int UTAPS = 0>
__global__ void __launch_bounds__(kPairThreads, MINB) ingest_resample_pair_kernel(const In* __restrict__ pcm, long long n_frames,
                                                                                   ResamplePlan plan, PairGeometry geo,
                                                                                   const __grid_constant__ UniformTaps uni,
                                                                                   const int* __restrict__ tap_first,
                                                                                   const float* __restrict__ tap_w, float* __restrict__ out,
                                                                                   long long n_real, long long out_len, long long n_items,
                                                                                   bool out_vec) {
    constexpr int bpf = static_cast<int>(sizeof(In)) * CH;          // bytes per frame: 2, 4 or 8
    constexpr int fpc = 16 / bpf;                                   // frames per 16-byte chunk
    extern __shared__ float4 smem4[];
    float* const span = reinterpret_cast<float*>(smem4);             // [rounds][sub_floats]
    uint32_t span_addr = static_cast<uint32_t>(__cvta_generic_to_shared(span));
    asm volatile("mov.u32 %0, %0;\n" : "+r"(span_addr));            // opaque: keeps the address in a register (ptxas otherwise
                                                                     // recomputes it from S2R CgaCtaId at every use)
    const uint32_t raw_addr = span_addr + static_cast<uint32_t>(geo.rounds * geo.sub_floats) * 4u;   // [kPairDepth][n_chunks] x 16 bytes
    const int nt = blockDim.x, t = threadIdx.x;
    const int per_item = G * nt * geo.rounds;
    const long long frames_item = static_cast<long long>(geo.round_stride) * geo.rounds;
    if (static_cast<long long>(blockIdx.x) >= n_items) return;
    const int items_mine = static_cast<int>((n_items - blockIdx.x + gridDim.x - 1) / gridDim.x);
    const char* const bytes = reinterpret_cast<const char*>(pcm);

    // first input frame of this block's n-th item; consecutive items of the block are ff_step frames apart
    const long long ff_step = frames_item * gridDim.x;
    const long long ff0 = frames_item * blockIdx.x + plan.first0 - plan.width - geo.lead;
    auto issue = [&](long long ff, int slot) {                       // raw chunks of the item that starts at frame ff -> ring slot
        const long long a_lo = ff & ~static_cast<long long>(fpc - 1);
        const uint32_t dst = raw_addr + static_cast<uint32_t>(slot * geo.n_chunks) * 16u;
        if (a_lo >= 0 && a_lo + static_cast<long long>(fpc) * geo.n_chunks <= n_frames) {          // wholly inside the stream
            const char* src = bytes + a_lo * bpf;
            for (int i = t; i < geo.n_chunks; i += nt) cp_async16_zfill(dst + static_cast<uint32_t>(i) * 16u, src + i * 16, 16);
        } else {
            for (int i = t; i < geo.n_chunks; i += nt) {
                const long long f = a_lo + static_cast<long long>(fpc) * i;
                long long nb = f < 0 ? 0 : (n_frames - f) * bpf;    // bytes of the stream from frame f on
                nb = nb < 0 ? 0 : nb > 16 ? 16 : nb;
                cp_async16_zfill(dst + static_cast<uint32_t>(i) * 16u, nb > 0 ? bytes + f * bpf : bytes, static_cast<int>(nb));
            }
        }
    };
#pragma unroll
    for (int d = 0; d < kPairDepth; ++d) {
        if (d < items_mine) issue(ff0 + d * ff_step, d);
        cp_async_commit();
    }

    // this thread's window inside a sub-span and its G shifted tap sets
    constexpr int kRegTaps = UORIG > 0 ? 1 : TE;                    // the few-phase form keeps no taps in registers
    float W[G][kRegTaps];
    int base;
    if constexpr (UORIG > 0) {
        base = t * ((G / UNEW) * UORIG);                            // a multiple of 4
    } else {
        int pos[G], ph[G];
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const int m = (G * t + g) / plan.new_f;
            ph[g] = (G * t + g) - m * plan.new_f;
            pos[g] = geo.lead + m * plan.orig_f + tap_first[ph[g]] - plan.first0;
        }
        base = pos[0] & ~3;
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const int sh = pos[g] - base;
            const float* w = tap_w + static_cast<size_t>(ph[g]) * plan.max_taps;
#pragma unroll
            for (int i = 0; i < TE; ++i) {
                const int k = i - sh;
                W[g][i] = (k >= 0 && k < plan.max_taps) ? w[k] : 0.f;
            }
        }
    }
    const uint32_t win_addr = span_addr + static_cast<uint32_t>(base) * 4u;

    int slot = 0;
    long long ff = ff0;                                              // first frame of item n
    long long j_item = static_cast<long long>(blockIdx.x) * per_item + G * t;   // this thread's first output of item n
    const long long j_step = static_cast<long long>(per_item) * gridDim.x;
    for (int n = 0; n < items_mine; ++n, ff += ff_step, j_item += j_step) {
        cp_async_wait<kPairDepth - 1>();                            // this thread's chunks of item n have landed
        __syncthreads();                                            // ... and everyone's; the previous filter pass is over
        {
            // raw -> float span, one sub-span per round: frames [i * round_stride, + sub_floats) of the item (the last
            // `overlap` of them are converted again as the head of the next round).  A thread keeps its offset k and
            // walks the rounds: two address increments per frame.
            const int skip = static_cast<int>(ff & static_cast<long long>(fpc - 1));
            const uint32_t src0 = raw_addr + static_cast<uint32_t>(slot * geo.n_chunks) * 16u + static_cast<uint32_t>(skip) * bpf;
            const uint32_t src_step = static_cast<uint32_t>(geo.round_stride) * bpf, dst_step = static_cast<uint32_t>(geo.sub_floats) * 4u;
            if constexpr (UORIG > 0) {
                // the few-phase kernels run one or two rounds per item (their stride is always even): batch over the
                // columns instead -- four loads in flight, then four stores, per round
                for (int i = 0; i < geo.rounds; ++i) {
                    const uint32_t src_i = src0 + static_cast<uint32_t>(i) * src_step, dst_i = span_addr + static_cast<uint32_t>(i) * dst_step;
                    for (int k = 2 * t; k < geo.sub_floats; k += 8 * nt) {
                        float2 v[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int kk = k + u * 2 * nt;
                            if (kk < geo.sub_floats) v[u] = raw_two_frames_to_mono<In, CH>(src_i + static_cast<uint32_t>(kk) * bpf);
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int kk = k + u * 2 * nt;
                            if (kk < geo.sub_floats) st_shared_f32x2(dst_i + static_cast<uint32_t>(kk) * 4u, v[u]);
                        }
                    }
                }
            } else if (geo.two) {
                // even stride, even first frame: two frames per step on 8-byte (two-frame) accesses; batched over the rounds
                for (int k = 2 * t; k < geo.sub_floats; k += 2 * nt) {
                    uint32_t src = src0 + static_cast<uint32_t>(k) * bpf, dst = span_addr + static_cast<uint32_t>(k) * 4u;
                    int i = 0;
                    for (; i + 4 <= geo.rounds; i += 4) {           // four loads in flight, then four stores
                        float2 v[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) v[u] = raw_two_frames_to_mono<In, CH>(src + static_cast<uint32_t>(u) * src_step);
#pragma unroll
                        for (int u = 0; u < 4; ++u) st_shared_f32x2(dst + static_cast<uint32_t>(u) * dst_step, v[u]);
                        src += 4u * src_step;
                        dst += 4u * dst_step;
                    }
                    for (; i < geo.rounds; ++i) {
                        st_shared_f32x2(dst, raw_two_frames_to_mono<In, CH>(src));
                        src += src_step;
                        dst += dst_step;
                    }
                }
            } else {
                for (int k = t; k < geo.sub_floats; k += nt) {
                    uint32_t src = src0 + static_cast<uint32_t>(k) * bpf, dst = span_addr + static_cast<uint32_t>(k) * 4u;
                    int i = 0;
                    for (; i + 4 <= geo.rounds; i += 4) {
                        float v[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) v[u] = raw_frame_to_mono<In, CH>(src + static_cast<uint32_t>(u) * src_step);
#pragma unroll
                        for (int u = 0; u < 4; ++u) st_shared_f32(dst + static_cast<uint32_t>(u) * dst_step, v[u]);
                        src += 4u * src_step;
                        dst += 4u * dst_step;
                    }
                    for (; i < geo.rounds; ++i) {
                        st_shared_f32(dst, raw_frame_to_mono<In, CH>(src));
                        src += src_step;
                        dst += dst_step;
                    }
                }
            }
        }
        __syncthreads();
        if (n + kPairDepth < items_mine) issue(ff + kPairDepth * ff_step, slot);
        cp_async_commit();
        slot = slot + 1 == kPairDepth ? 0 : slot + 1;

        const bool whole = out_vec && j_item - G * t + per_item <= n_real;   // no padding, no tail inside this item
        float* o = out + j_item;
        uint32_t xa = win_addr;
#pragma unroll 2
        for (int i = 0; i < geo.rounds; ++i) {
            float acc[G];
#pragma unroll
            for (int g = 0; g < G; ++g) acc[g] = 0.f;
#pragma unroll
            for (int c = 0; c < TE / 4; ++c) {
                const float4 v = ld_shared_f32x4(xa + 16u * c);
                const float x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int e = 0; e < 4; ++e)
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        if constexpr (UORIG > 0) {
                            constexpr int kDummy = 0;
                            (void)kDummy;
                            const int k = 4 * c + e - (g / UNEW) * UORIG;             // compile-time after unrolling
                            if (k >= 0 && k < UTAPS) acc[g] = fmaf(uni.w[(g % UNEW) * UTAPS + k], x[e], acc[g]);
                        } else {
                            acc[g] = fmaf(W[g][4 * c + e], x[e], acc[g]);
                        }
                    }
            }
            if (whole) {
                if constexpr (G % 4 == 0) {
#pragma unroll
                    for (int g = 0; g < G; g += 4) *reinterpret_cast<float4*>(o + g) = make_float4(acc[g], acc[g + 1], acc[g + 2], acc[g + 3]);
                } else if constexpr (G == 2) {
                    *reinterpret_cast<float2*>(o) = make_float2(acc[0], acc[1]);
                } else {
                    o[0] = acc[0];
                }
            } else {
                const long long j = j_item + static_cast<long long>(G) * nt * i;
#pragma unroll
                for (int g = 0; g < G; ++g)
                    if (j + g < out_len) o[g] = j + g < n_real ? acc[g] : 0.f;         // IR:150-154
            }
            o += G * nt;
            xa += static_cast<uint32_t>(geo.sub_floats) * 4u;
        }
    }
    cp_async_wait<0>();
}

template <typename In, int CH, int TE, int G, int MINB = 2, int UORIG = 0, int UNEW = 1, int UTAPS = 0>
cudaError_t launch_pair(const In* pcm, long long n_frames, const ResamplePlan& plan, const PairGeometry& geo, const UniformTaps& uni,
                        const int* tap_first, const float* tap_w, float* out, long long n_real, long long out_len, int threads,
                        size_t smem, int sms, cudaStream_t stream) {
    const long long per_item = static_cast<long long>(G) * threads * geo.rounds;
    const long long n_items = (out_len + per_item - 1) / per_item;
    auto kernel = ingest_resample_pair_kernel<In, CH, TE, G, MINB, UORIG, UNEW, UTAPS>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    const long long cap = static_cast<long long>(sms) * MINB;          // resident blocks per SM
    const unsigned grid = static_cast<unsigned>(n_items < cap ? n_items : cap);
    const bool out_vec = (reinterpret_cast<uintptr_t>(out) & (G % 4 == 0 ? 15 : 7)) == 0;
    kernel<<<grid, threads, smem, stream>>>(pcm, n_frames, plan, geo, uni, tap_first, tap_w, out, n_real, out_len, n_items, out_vec);
    return cudaGetLastError();
}

// The pair kernel applies when a block of <= 320 threads holds whole periods of the phase pattern and the stream is mono
// or stereo on a 16-byte aligned base.  Two outputs per thread when their window (alignment slack 3 + pair_shift_max +
// taps) fits 28 floats: 44.1 / 48 kHz and every rate below.  ONE output per thread on a 40-float window for the
// 34-37-tap rates (88.2 / 96 kHz): two tap sets of 44 need 158 registers, one block per SM, and measured 0.81x of
// the one-output instantiation (2.06 against 2.54 TB/s at 96 kHz): occupancy carries the kernel.  SAD_INGEST_PAIR=0 turns it off.
template <typename In>
bool try_pair(const In* pcm, long long n_frames, int channels, const ResamplePlan& plan, const IngestTables& tb, float* out,
              long long n_real, long long out_len, cudaStream_t stream, cudaError_t* err) {
    static const bool enabled = [] { const char* v = getenv("SAD_INGEST_PAIR"); return !(v && v[0] == '0'); }();
    static const bool few_phase = [] { const char* v = getenv("SAD_INGEST_UNIFORM"); return !(v && v[0] == '0'); }();
    if (!enabled || channels > 2 || n_frames < 1 || (reinterpret_cast<uintptr_t>(pcm) & 15) != 0) return false;
    int dev = 0, sms = 0;
    if ((*err = cudaGetDevice(&dev)) != cudaSuccess || (*err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess)
        return true;
    PairChoice c{};
    if (!choose_pair_geometry(plan, static_cast<int>(sizeof(In)) * channels, out_len, sms, &c, few_phase ? tb.uniform : nullptr)) return false;
    const PairGeometry& geo = c.geo;
    const int threads = c.threads;
    const size_t smem = c.smem;
    const int* tap_first = tb.tap_first;
    const float* tap_w = tb.tap_w;
    static const UniformTaps no_taps{};
#define SAD_PAIR_CASE(TE_, G_)                                                                                                    \
    *err = channels == 2 ? launch_pair<In, 2, TE_, G_>(pcm, n_frames, plan, geo, no_taps, tap_first, tap_w, out, n_real, out_len, \
                                                       threads, smem, sms, stream)                                                \
                         : launch_pair<In, 1, TE_, G_>(pcm, n_frames, plan, geo, no_taps, tap_first, tap_w, out, n_real, out_len, \
                                                       threads, smem, sms, stream)
#define SAD_UNIFORM_CASE(ORIG_, NEW_, G_, TAPS_)                                                                                  \
    if (u.orig == ORIG_ && u.phases == NEW_ && u.outputs == G_ && u.taps == TAPS_) {                                             \
        constexpr int kWin = ((G_ / NEW_ - 1) * ORIG_ + TAPS_ + 3) / 4 * 4;                                                       \
        *err = channels == 2 ? launch_pair<In, 2, kWin, G_, 4, ORIG_, NEW_, TAPS_>(pcm, n_frames, plan, geo, u, tap_first, tap_w, \
                                                                                   out, n_real, out_len, threads, smem, sms,      \
                                                                                   stream)                                        \
                             : launch_pair<In, 1, kWin, G_, 4, ORIG_, NEW_, TAPS_>(pcm, n_frames, plan, geo, u, tap_first, tap_w, \
                                                                                   out, n_real, out_len, threads, smem, sms,      \
                                                                                   stream);                                       \
        return true;                                                                                                              \
    }
    if (c.uniform) {
        const UniformTaps& u = *tb.uniform;                          // the rows of kUniformConfigs
        SAD_UNIFORM_CASE(3, 1, 4, 40)
        SAD_UNIFORM_CASE(2, 1, 4, 28)
        SAD_UNIFORM_CASE(6, 1, 2, 76)
        SAD_UNIFORM_CASE(3, 2, 8, 24)
        SAD_UNIFORM_CASE(1, 2, 8, 16)
        *err = cudaErrorInvalidValue;                                // a config without an instantiation
        return true;
    }
    switch (c.window) {
        case 20: SAD_PAIR_CASE(20, 2); break;
        case 24: SAD_PAIR_CASE(24, 2); break;
        case 28: SAD_PAIR_CASE(28, 2); break;
        default: SAD_PAIR_CASE(40, 1); break;
    }
#undef SAD_PAIR_CASE
#undef SAD_UNIFORM_CASE
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
        const bool pair = sample_format == 0
            ? try_pair(static_cast<const int16_t*>(pcm), n_frames, channels, *plan, tb, out, n_real, out_len, stream, &e)
            : try_pair(static_cast<const float*>(pcm), n_frames, channels, *plan, tb, out, n_real, out_len, stream, &e);
        if (pair) {
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
