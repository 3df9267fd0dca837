// K4 / K5: global average pool + per-head MLP (BatchNorm1d folded) + merge + sigmoid/threshold decision,
// and the per-clip mean.  fp32 throughout (the tolerance budget is spent on the bf16 convolutions).
// Replaces BinaryClassifier.head (reference modular/source/inference_runner.py:36-48, eval mode),
// ModularMultiHeadClassifier.forward's merge (:62-73), interpret_multihead_logits (:194-214) and the clip
// mean (:328-334).
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "head.h"

namespace sad {

namespace {

// grid (ceil(B / kSegPerCta), H), 256 threads.  feats: NHWC bf16 [H*B][256 px][512]; weights transposed [in][out] fp32.
// Each CTA serves kSegPerCta segments of one head.  Measured on B200 (768 head-segments per launch): 1 segment per
// CTA (768 CTAs, full occupancy) 3.2 ms/step, 4 per CTA 6.2 ms, 8 per CTA 7.9 ms -- the kernel is latency-bound, so
// thread-level parallelism beats re-using the 1.5 MB of weights that sit in L2 anyway.
constexpr int kSegPerCta = 1;

__global__ void __launch_bounds__(256) head_mlp_kernel(const __nv_bfloat16* __restrict__ feats, HeadWeights hw, int B,
                                                       float* __restrict__ head_logits) {
    __shared__ float pooled[kSegPerCta][512];
    __shared__ float h1[kSegPerCta][512];
    __shared__ float h2[kSegPerCta][256];
    __shared__ float red[kSegPerCta][2][8];
    const int b0 = blockIdx.x * kSegPerCta, h = blockIdx.y, t = threadIdx.x;
    const int nseg = min(kSegPerCta, B - b0);
    for (int s = 0; s < nseg; ++s) {            // global average pool over the 16x16 map, two channels per thread
        const size_t n = static_cast<size_t>(h) * B + b0 + s;
        const __nv_bfloat162* f = reinterpret_cast<const __nv_bfloat162*>(feats + n * 256 * 512);
        float s0 = 0.f, s1 = 0.f;
#pragma unroll 8
        for (int p = 0; p < 256; ++p) {
            const float2 v = __bfloat1622float2(f[p * 256 + t]);
            s0 += v.x;
            s1 += v.y;
        }
        pooled[s][2 * t] = s0 * (1.0f / 256.0f);
        pooled[s][2 * t + 1] = s1 * (1.0f / 256.0f);
    }
    for (int s = nseg; s < kSegPerCta; ++s) {
        pooled[s][2 * t] = 0.f;
        pooled[s][2 * t + 1] = 0.f;
    }
    __syncthreads();

    const float* w1 = hw.w1t + static_cast<size_t>(h) * 512 * 512;
    float a0[kSegPerCta], a1[kSegPerCta];
#pragma unroll
    for (int s = 0; s < kSegPerCta; ++s) {
        a0[s] = hw.b1[h * 512 + t];
        a1[s] = hw.b1[h * 512 + t + 256];
    }
#pragma unroll 2
    for (int i = 0; i < 512; ++i) {
        const float wa = __ldg(w1 + i * 512 + t), wb = __ldg(w1 + i * 512 + t + 256);
#pragma unroll
        for (int s = 0; s < kSegPerCta; ++s) {
            const float x = pooled[s][i];
            a0[s] = fmaf(x, wa, a0[s]);
            a1[s] = fmaf(x, wb, a1[s]);
        }
    }
#pragma unroll
    for (int s = 0; s < kSegPerCta; ++s) {
        h1[s][t] = fmaxf(a0[s], 0.f);
        h1[s][t + 256] = fmaxf(a1[s], 0.f);
    }
    __syncthreads();

    const float* w2 = hw.w2t + static_cast<size_t>(h) * 512 * 256;
    float c[kSegPerCta];
#pragma unroll
    for (int s = 0; s < kSegPerCta; ++s) c[s] = hw.b2[h * 256 + t];
#pragma unroll 2
    for (int i = 0; i < 512; ++i) {
        const float w = __ldg(w2 + i * 256 + t);
#pragma unroll
        for (int s = 0; s < kSegPerCta; ++s) c[s] = fmaf(h1[s][i], w, c[s]);
    }
#pragma unroll
    for (int s = 0; s < kSegPerCta; ++s) h2[s][t] = fmaxf(c[s], 0.f);
    __syncthreads();

    const float* w3 = hw.w3 + static_cast<size_t>(h) * 2 * 256;   // [2][256] as in nn.Linear
    const float w30 = w3[t], w31 = w3[256 + t];
#pragma unroll
    for (int s = 0; s < kSegPerCta; ++s) {
        float z0 = h2[s][t] * w30, z1 = h2[s][t] * w31;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            z0 += __shfl_xor_sync(0xffffffffu, z0, o);
            z1 += __shfl_xor_sync(0xffffffffu, z1, o);
        }
        if ((t & 31) == 0) {
            red[s][0][t >> 5] = z0;
            red[s][1][t >> 5] = z1;
        }
    }
    __syncthreads();
    if (t < 2 * kSegPerCta) {
        const int s = t >> 1, j = t & 1;
        if (s < nseg) {
            float z = hw.b3[h * 2 + j];
            for (int i = 0; i < 8; ++i) z += red[s][j][i];
            head_logits[(static_cast<size_t>(h) * B + b0 + s) * 2 + j] = z;   // index 0 = Real, 1 = Synthetic
        }
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

cudaError_t head_mlp_launch(const __nv_bfloat16* feats, const HeadWeights& hw, int B, int H, float* head_logits,
                            cudaStream_t stream, long long* launches) {
    head_mlp_kernel<<<dim3((B + kSegPerCta - 1) / kSegPerCta, H), 256, 0, stream>>>(feats, hw, B, head_logits);
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
