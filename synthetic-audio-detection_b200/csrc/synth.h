// Launch wrapper of the synthetic-corpus generator (synth.cu): bench / corpus support, not part of the reference path.
#pragma once
#include <cuda_runtime.h>

namespace sad {

// out [n][128000] fp32: segment `first + i` of the SURVEY 8(d) noise/tone corpus for base seed `seed`.  Counter
// based: every sample is a pure function of (seed, global segment index, sample index), so any shard, chunk or rank
// produces the same bytes for the same global segment.
cudaError_t synth_segments_launch(float* out, long long first, int n, unsigned long long seed, cudaStream_t stream,
                                  long long* launches);

}  // namespace sad
