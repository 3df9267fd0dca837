// Ingest stage (ingest.cu): interleaved PCM -> mono float32 at 32 kHz, zero-padded to one window.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

namespace sad {

constexpr int kIngestRate = 32000;          // IR:258 sample_rate
constexpr long long kIngestWindow = 128000; // int(4.0 * 32000), IR:150

struct ResamplePlan {
    int orig_f, new_f;   // sample rates divided by their gcd
    int width;           // torchaudio's one-sided kernel reach in input frames
    int taps_full;       // 2*width + orig_f: length of torchaudio's dense kernel
    int max_taps;        // taps kept per phase (the band where the Hann window argument is not clamped)
};

struct IngestTables {
    const int* tap_first;    // [new_f] first kept tap of each phase
    const float* tap_w;      // [new_f][max_taps]
};

// Output length of the reference's preprocess_waveform (resampled length with torchaudio's float32 ceil, at least one
// window); *n_real = samples before padding.  < 0 on bad arguments.
long long ingest_length(long long n_frames, int sr_in, long long* n_real);
// Per-phase tap bands of torchaudio's sinc_interp_hann kernel for sr_in -> 32 kHz.  false: ratio not supported.
bool build_resample_taps(int sr_in, ResamplePlan* plan, std::vector<int>* first, std::vector<float>* weights);
size_t ingest_smem_bytes(const ResamplePlan& plan);
// plan == nullptr: sr_in is already 32 kHz (mix + pad only).  sample_format 0 = int16, 1 = float32.
cudaError_t ingest_launch(const void* pcm, int sample_format, long long n_frames, int channels, const ResamplePlan* plan,
                          const IngestTables& tb, float* out, long long n_real, long long out_len, cudaStream_t stream,
                          long long* launches);

}  // namespace sad
