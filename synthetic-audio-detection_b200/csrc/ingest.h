// Ingest stage (ingest.cu): interleaved PCM -> mono float32 at 32 kHz, zero-padded to one window.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#include "ingest_taps.h"

namespace sad {

struct IngestTables {
    const int* tap_first;    // [new_f] first kept tap of each phase
    const float* tap_w;      // [new_f][max_taps]
    const UniformTaps* uniform = nullptr;   // HOST pointer: the taps of a few-phase ratio as kernel parameters, or null
};

size_t ingest_smem_bytes(const ResamplePlan& plan);
// plan == nullptr: sr_in is already 32 kHz (mix + pad only).  sample_format 0 = int16, 1 = float32.
cudaError_t ingest_launch(const void* pcm, int sample_format, long long n_frames, int channels, const ResamplePlan* plan,
                          const IngestTables& tb, float* out, long long n_real, long long out_len, cudaStream_t stream,
                          long long* launches);

}  // namespace sad
