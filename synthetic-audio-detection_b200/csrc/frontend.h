// Launch wrappers of the front-end kernels (frontend.cu) and the small device-resident tables they read.
#pragma once
#include <cstddef>
#include <cstdint>
#include "act.cuh"
#include <cuda_runtime.h>

namespace sad {

// Banded form of the [1025,128] mel filterbank: filter m covers bins start[m] .. start[m]+count[m]-1 with
// weights w[off[m] ...].  Only bins <= 768 may carry weight (f_max = 12 kHz), checked when it is built.
struct MelTable {
    int n_weights;
    int start[128];
    int count[128];
    int off[128];
    float w[1536];
    // the same bands widened to 4-bin boundaries with zero weights (logmel_kernel reads taps and weights with 16-byte
    // loads): start4 = start & ~3, count4 a multiple of 4, weights at w4[off4 ...] (off4 a multiple of 4)
    int start4[128];
    int count4[128];
    int off4[128];
    float w4[2816];
};

// Two-tap anti-aliased bilinear weights (ATen upsample_bilinear2d_aa), 251 -> 512 columns, 128 -> 512 rows.
struct ResizeTable {
    int w_idx[512];
    float w_w[1024];
    int h_idx[512];
    float h_w[1024];
};

size_t stft_smem_bytes();

// db_work [B][128][251] receives the final (clamped) log-mel dB, out_db (optional) a copy, mu_sigma [B][2] the mean and
// unbiased std.  scratch [B][251][128]: frame-major unclamped dB of the one-launch kernel (nullptr or SAD_FE_V1=1: the
// round-1 three-launch path, which needs segmax [B]).
cudaError_t frontend_logmel_launch(const float* pcm, int B, const float* window, const MelTable* mel, float* db_work,
                                   unsigned* segmax, float* out_db, float* mu_sigma, float* scratch, cudaStream_t stream,
                                   long long* launches);
cudaError_t image_launch_f32(const float* db, const float* mu_sigma, const ResizeTable* rt, float* img, int B,
                             cudaStream_t stream, long long* launches);
cudaError_t image_launch_bf16(const float* db, const float* mu_sigma, const ResizeTable* rt, act_t* img, int B,
                              cudaStream_t stream, long long* launches);
cudaError_t im2col_stem3_launch(const float* x, act_t* A, int B, cudaStream_t stream, long long* launches);
cudaError_t maxpool_launch(const act_t* in, act_t* out, long long n_img, cudaStream_t stream,
                           long long* launches);
cudaError_t slice_gate_launch(const float* wf, long long n_windows, long long window, long long hop, float thr,
                              uint8_t* keep, cudaStream_t stream, long long* launches);
cudaError_t gather_windows_launch(const float* wf, const long long* starts, int n_kept, long long window, float* dst,
                                  cudaStream_t stream, long long* launches);

}  // namespace sad
