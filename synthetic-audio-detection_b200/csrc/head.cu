// K4 / K5: global average pool + per-head MLP (BatchNorm1d folded) + merge + sigmoid/threshold decision,
// and the per-clip mean.  fp32 throughout (the tolerance budget is spent on the bf16 convolutions).
// Replaces BinaryClassifier.head (reference modular/source/inference_runner.py:36-48, eval mode),
// ModularMultiHeadClassifier.forward's merge (:62-73), interpret_multihead_logits (:194-214) and the clip
// mean (:328-334).
#include <cuda_runtime.h>

#include <cstdint>

#include "head.h"

namespace sad {

namespace {

// grid (ceil(B / kSegs), H), 256 threads: one CTA per (head, group of kSegs segments).  Every weight a thread loads is
// used for all kSegs segments, so the L2 -> SM weight stream is 1.5 MB per kSegs segments instead of per segment
// (ncu, round 2: with one segment per CTA the kernel moves 1.15 GB of weights per chunk out of L2 and takes 105 us
// against a 31 us floor for reading the features).  Per segment the arithmetic (K split, summation order) does not
// depend on kSegs, so results are bit-identical for every grouping.  MEASURED: grouping 2 or 4 segments is SLOWER
// (fewer, longer CTAs: the pooling and GEMV phases of a CTA are serial and nothing overlaps them), so kSegs = 1 ships;
// the switch stays for the next attempt (more warps per CTA, or the pooling fused into layer4's epilogue).
// feats: NHWC [H*B][256 px][F] (act_t); weights transposed [in][out] fp32 so a thread reads 4 consecutive outputs with
// one 16-byte load; the K dimension is split over thread groups and reduced through shared memory.
// F = trunk feature width: 512 (resnet18/34) or 2048 (Bottleneck nets).
#ifndef SAD_HEAD_SEGS
#define SAD_HEAD_SEGS 1   // measured on B200 (6 heads, chunk 128): 1 -> 1.97 ms per step, 2 -> 2.43, 4 -> 2.34
#endif
__host__ __device__ constexpr int head_segs(int F) { return F <= 512 ? SAD_HEAD_SEGS : (SAD_HEAD_SEGS > 2 ? 2 : SAD_HEAD_SEGS); }   // segments per CTA (static shared memory <= 48 KB)

template <int F>
__global__ void __launch_bounds__(256) head_mlp_kernel(const act_t* __restrict__ feats, HeadWeights hw, int B,
                                                       float* __restrict__ head_logits) {
    constexpr int kC8 = F / 8;            // 16-byte channel groups per pixel
    constexpr int kG = 256 / kC8;         // pixel groups in the pooling phase (4 for F=512, 1 for F=2048)
    constexpr int kPx = 256 / kG;         // pixels per group
    constexpr int kSegs = head_segs(F);
    __shared__ __align__(16) float part_flat[kSegs][2048];   // partial sums (pool: kG pixel groups x F; layers: K splits x 512)
    __shared__ __align__(16) float act_buf[kSegs][F];        // pooled features, then (after a barrier each) h1 and h2
    float (*pooled)[F] = act_buf;
    float (*h1)[F] = act_buf;                                // first 512 entries of a row
    float (*h2)[F] = act_buf;                                // first 256 entries of a row
    __shared__ float red[kSegs][2][8];
    const int b0 = blockIdx.x * kSegs, h = blockIdx.y, t = threadIdx.x;
    const int nseg = B - b0 < kSegs ? B - b0 : kSegs;

    // ---- global average pool over the 16x16 map: thread = (pixel group g of 64 px, 8 channels c8)
    for (int sg = 0; sg < nseg; ++sg) {
        const size_t n = static_cast<size_t>(h) * B + b0 + sg;
        const int c8 = t % kC8, g = t / kC8;
        const uint4* f = reinterpret_cast<const uint4*>(feats + n * 256 * F) + c8;   // kC8 uint4 per pixel
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
        for (int p = 0; p < kPx; ++p) {
            const uint4 v = __ldg(f + (g * kPx + p) * kC8);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                acc[2 * j] += act_lo(w[j]);
                acc[2 * j + 1] += act_hi(w[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) part_flat[sg][g * F + c8 * 8 + j] = acc[j];
    }
    __syncthreads();
    for (int sg = 0; sg < kSegs; ++sg)
        for (int c = t; c < F; c += 256) {
            float sum = 0.f;
            if (sg < nseg) {
                sum = part_flat[sg][c];
#pragma unroll
                for (int g = 1; g < kG; ++g) sum += part_flat[sg][g * F + c];
            }
            pooled[sg][c] = sum * (1.0f / 256.0f);          // missing segments of a ragged tail run on zeros
        }
    __syncthreads();

    // ---- Linear(F,512)+BN folded, ReLU: thread = (K half kh, 4 outputs o4)
    {
        const int o4 = t & 127, kh = t >> 7;
        const float4* w1 = reinterpret_cast<const float4*>(hw.w1t + static_cast<size_t>(h) * F * 512) + o4;
        float4 a[kSegs];
#pragma unroll
        for (int sg = 0; sg < kSegs; ++sg) a[sg] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
        for (int i = kh * (F / 2); i < (kh + 1) * (F / 2); ++i) {
            const float4 w = __ldg(w1 + i * 128);
#pragma unroll
            for (int sg = 0; sg < kSegs; ++sg) {
                const float x = pooled[sg][i];
                a[sg].x = fmaf(x, w.x, a[sg].x); a[sg].y = fmaf(x, w.y, a[sg].y);
                a[sg].z = fmaf(x, w.z, a[sg].z); a[sg].w = fmaf(x, w.w, a[sg].w);
            }
        }
#pragma unroll
        for (int sg = 0; sg < kSegs; ++sg) *reinterpret_cast<float4*>(&part_flat[sg][kh * 512 + o4 * 4]) = a[sg];
    }
    __syncthreads();
    for (int sg = 0; sg < kSegs; ++sg)
        for (int c = t; c < 512; c += 256)
            h1[sg][c] = fmaxf(part_flat[sg][c] + part_flat[sg][512 + c] + hw.b1[h * 512 + c], 0.f);
    __syncthreads();

    // ---- Linear(512,256)+BN folded, ReLU: thread = (K quarter kq, 4 outputs o4)
    {
        const int o4 = t & 63, kq = t >> 6;
        const float4* w2 = reinterpret_cast<const float4*>(hw.w2t + static_cast<size_t>(h) * 512 * 256) + o4;
        float4 a[kSegs];
#pragma unroll
        for (int sg = 0; sg < kSegs; ++sg) a[sg] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
        for (int i = kq * 128; i < kq * 128 + 128; ++i) {
            const float4 w = __ldg(w2 + i * 64);
#pragma unroll
            for (int sg = 0; sg < kSegs; ++sg) {
                const float x = h1[sg][i];
                a[sg].x = fmaf(x, w.x, a[sg].x); a[sg].y = fmaf(x, w.y, a[sg].y);
                a[sg].z = fmaf(x, w.z, a[sg].z); a[sg].w = fmaf(x, w.w, a[sg].w);
            }
        }
#pragma unroll
        for (int sg = 0; sg < kSegs; ++sg) *reinterpret_cast<float4*>(&part_flat[sg][kq * 512 + o4 * 4]) = a[sg];
    }
    __syncthreads();
    for (int sg = 0; sg < kSegs; ++sg)
        h2[sg][t] = fmaxf(((part_flat[sg][t] + part_flat[sg][512 + t]) + part_flat[sg][1024 + t]) + part_flat[sg][1536 + t] +
                          hw.b2[h * 256 + t], 0.f);
    __syncthreads();

    // ---- Linear(256,2): block reduction
    const float* w3 = hw.w3 + static_cast<size_t>(h) * 2 * 256;   // [2][256] as in nn.Linear
    const float w30 = w3[t], w31 = w3[256 + t];
#pragma unroll
    for (int sg = 0; sg < kSegs; ++sg) {
        float z0 = h2[sg][t] * w30, z1 = h2[sg][t] * w31;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            z0 += __shfl_xor_sync(0xffffffffu, z0, o);
            z1 += __shfl_xor_sync(0xffffffffu, z1, o);
        }
        if ((t & 31) == 0) {
            red[sg][0][t >> 5] = z0;
            red[sg][1][t >> 5] = z1;
        }
    }
    __syncthreads();
    if (t < 2 * nseg) {
        const int sg = t >> 1, k = t & 1;
        float z = hw.b3[h * 2 + k];
        for (int i = 0; i < 8; ++i) z += red[sg][k][i];
        head_logits[(static_cast<size_t>(h) * B + b0 + sg) * 2 + k] = z;   // index 0 = Real, 1 = Synthetic
    }
}

// One warp per segment; lane i holds head i's (real, synthetic) logits.  N <= 31.
__global__ void __launch_bounds__(256) merge_decide_kernel(const float* __restrict__ head_logits, int B, int N, float thr,
                                                           float* __restrict__ logits, float* __restrict__ probs,
                                                           int* __restrict__ labels) {
    const int b = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= B) return;
    float real = 0.f, syn = 0.f;
    if (lane < N) {
        real = head_logits[(static_cast<size_t>(lane) * B + b) * 2];
        syn = head_logits[(static_cast<size_t>(lane) * B + b) * 2 + 1];
    }
    float acc = 0.f;
    for (int i = 0; i < N; ++i) acc += __shfl_sync(0xffffffffu, real, i);   // sequential order, like torch.mean
    const float real_mean = acc / static_cast<float>(N);
    const float z = lane < N ? syn : real_mean;                               // column `lane` of [syn.., real_mean]
    const float s = 1.0f / (1.0f + expf(-z));
    if (lane <= N) {
        if (logits) logits[static_cast<size_t>(b) * (N + 1) + lane] = z;
        if (probs) probs[static_cast<size_t>(b) * (N + 1) + lane] = s;
    }
    // decision: gather the N+1 probabilities into every lane's registers is overkill; use ballots
    const unsigned below = __ballot_sync(0xffffffffu, lane < N && s < thr);
    const bool all_below = below == ((N >= 32) ? 0xffffffffu : ((1u << N) - 1u));
    const float s_real = __shfl_sync(0xffffffffu, s, N);
    float best = lane < N ? s : -1.f;
    int arg = lane < N ? lane : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
        if (ob > best || (ob == best && oa < arg)) {
            best = ob;
            arg = oa;
        }
    }
    if (lane == 0 && labels) labels[b] = (s_real >= thr && all_below) ? N : arg;
}

// One warp per clip over SORTED clip ids: binary-search the run, sum rows sequentially per column (numpy's
// order for np.mean(axis=0)), divide, decide.
__global__ void __launch_bounds__(256) clip_reduce_kernel(const float* __restrict__ probs, const int* __restrict__ clip_id,
                                                          int B, int n_clips, int N, float thr,
                                                          float* __restrict__ clip_probs, int* __restrict__ clip_label) {
    const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (c >= n_clips) return;
    int lo = 0, hi = B;
    while (lo < hi) {   // first index with clip_id >= c
        const int mid = (lo + hi) >> 1;
        if (clip_id[mid] < c) lo = mid + 1; else hi = mid;
    }
    const int first = lo;
    hi = B;
    while (lo < hi) {   // first index with clip_id > c
        const int mid = (lo + hi) >> 1;
        if (clip_id[mid] <= c) lo = mid + 1; else hi = mid;
    }
    const int last = lo;
    const int cnt = last - first;
    float acc = 0.f;
    if (lane <= N)
        for (int r = first; r < last; ++r) acc += probs[static_cast<size_t>(r) * (N + 1) + lane];
    const float mean = cnt > 0 ? acc / static_cast<float>(cnt) : 0.f;
    if (lane <= N) clip_probs[static_cast<size_t>(c) * (N + 1) + lane] = mean;
    const unsigned below = __ballot_sync(0xffffffffu, lane < N && mean < thr);
    const bool all_below = below == ((1u << N) - 1u);
    const float s_real = __shfl_sync(0xffffffffu, mean, N);
    float best = lane < N ? mean : -1.f;
    int arg = lane < N ? lane : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
        if (ob > best || (ob == best && oa < arg)) {
            best = ob;
            arg = oa;
        }
    }
    if (lane == 0) clip_label[c] = cnt == 0 ? -1 : ((s_real >= thr && all_below) ? N : arg);
}

}  // namespace

cudaError_t head_mlp_launch(const act_t* feats, const HeadWeights& hw, int B, int H, int features, float* head_logits,
                            cudaStream_t stream, long long* launches) {
    if (features == 512) head_mlp_kernel<512><<<dim3((B + head_segs(512) - 1) / head_segs(512), H), 256, 0, stream>>>(feats, hw, B, head_logits);
    else if (features == 2048) head_mlp_kernel<2048><<<dim3((B + head_segs(2048) - 1) / head_segs(2048), H), 256, 0, stream>>>(feats, hw, B, head_logits);
    else return cudaErrorInvalidValue;
    if (launches) *launches += 1;
    return cudaGetLastError();
}
cudaError_t merge_decide_launch(const float* head_logits, int B, int N, float thr, float* logits, float* probs, int* labels,
                                cudaStream_t stream, long long* launches) {
    merge_decide_kernel<<<(B + 7) / 8, 256, 0, stream>>>(head_logits, B, N, thr, logits, probs, labels);
    if (launches) *launches += 1;
    return cudaGetLastError();
}
cudaError_t clip_reduce_launch(const float* probs, const int* clip_id, int B, int n_clips, int N, float thr,
                               float* clip_probs, int* clip_label, cudaStream_t stream, long long* launches) {
    if (n_clips <= 0) return cudaSuccess;
    clip_reduce_kernel<<<(n_clips + 7) / 8, 256, 0, stream>>>(probs, clip_id, B, n_clips, N, thr, clip_probs, clip_label);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

}  // namespace sad
