// Launch wrappers of the head / merge / clip kernels (head.cu).
#pragma once
#include "act.cuh"
#include <cuda_runtime.h>

namespace sad {

// Folded head MLP of all H heads (device pointers, fp32):
//   w1t [H][F in][512 out],   b1 [H][512]   Linear(F,512)+BatchNorm1d(512) folded, transposed for coalescing; F = 512 | 2048
//   w2t [H][512 in][256 out], b2 [H][256]   Linear(512,256)+BatchNorm1d(256) folded
//   w3  [H][2][256],          b3 [H][2]     Linear(256,2): row 0 = Real, row 1 = Synthetic
struct HeadWeights {
    const float* w1t;
    const float* b1;
    const float* w2t;
    const float* b2;
    const float* w3;
    const float* b3;
};

cudaError_t head_mlp_launch(const act_t* feats, const HeadWeights& hw, int B, int H, int features, float* head_logits,
                            cudaStream_t stream, long long* launches);
cudaError_t merge_decide_launch(const float* head_logits, int B, int N, float thr, float* logits, float* probs, int* labels,
                                cudaStream_t stream, long long* launches);
cudaError_t clip_reduce_launch(const float* probs, const int* clip_id, int B, int n_clips, int N, float thr,
                               float* clip_probs, int* clip_label, cudaStream_t stream, long long* launches);

}  // namespace sad
