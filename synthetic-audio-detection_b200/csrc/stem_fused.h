// Launch description of the fused stem kernel (stem_fused.cu).
#pragma once
#include <cuda.h>
#include "act.cuh"
#include <cuda_runtime.h>

namespace sad {

struct alignas(64) StemLaunch {
    CUtensorMap w_map;      // stem weights [H*64 rows][64 k] bf16: k = ky*8+kx (kx<7), k=56..58 = bias hi/mid/lo; box {64,128}
    CUtensorMap out_map;    // pooled output viewed as {64 ch, pixels}; box {64, 128} = one pooled row of one head
    const act_t* img;   // [B][512][512] bf16 standardised, resized log-mel image
    int B, H;
    int G;                  // head pairs (filled by the launcher)
    int total_units;        // filled by the launcher
};

cudaError_t stem_fused_launch(const StemLaunch& p, int num_sms, cudaStream_t stream);

}  // namespace sad
