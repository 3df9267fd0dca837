// Host-side arithmetic of the ingest stage, free of CUDA so that the CPU test suite can build it with g++
// (csrc/ingest_host_check.cpp, tests/test_ingest_host.py): output length and the per-phase tap bands of torchaudio's
// sinc_interp_hann resampling kernel (torchaudio/functional/functional.py:1452-1577, defaults width 6, rolloff 0.99).
#pragma once
#include <cmath>
#include <cstddef>
#include <vector>

namespace sad {

constexpr int kIngestRate = 32000;          // IR:258 sample_rate
constexpr long long kIngestWindow = 128000; // int(4.0 * 32000), IR:150

struct ResamplePlan {
    int orig_f, new_f;   // sample rates divided by their gcd
    int width;           // torchaudio's one-sided kernel reach in input frames
    int taps_full;       // 2*width + orig_f: length of torchaudio's dense kernel
    int max_taps;        // taps kept per phase (the band where the Hann window argument is not clamped)
    // geometry of the bands, for the kernel that computes two adjacent outputs per thread (ingest.cu):
    int first0;          // first kept tap of phase 0
    int first_spread;    // max over phases of (first kept tap) - first0
    int pair_shift_max;  // max distance in input frames between the band starts of outputs j and j+1 (j even);
                         // -1 when band starts are not non-decreasing in j
};

// Output length of the reference's preprocess_waveform (resampled length with torchaudio's float32 ceil, at least one
// window); *n_real = samples before padding.  < 0 on bad arguments.
inline long long ingest_length(long long n_frames, int sr_in, long long* n_real_out) {
    if (n_frames < 0 || sr_in <= 0) return -1;
    long long n_real = n_frames;
    if (sr_in != kIngestRate) {
        long long a = sr_in, b = kIngestRate;
        while (b) { const long long t = a % b; a = b; b = t; }
        const long long orig = sr_in / a, nw = kIngestRate / a;
        // torch.ceil(torch.as_tensor(new * length / orig)): the quotient is a Python float (correctly rounded double) that
        // as_tensor stores as float32 BEFORE the ceil
        const float q = static_cast<float>(static_cast<double>(nw * n_frames) / static_cast<double>(orig));
        n_real = static_cast<long long>(std::ceil(q));
    }
    if (n_real_out) *n_real_out = n_real;
    return n_real < kIngestWindow ? kIngestWindow : n_real;
}

// torchaudio.functional._get_sinc_resample_kernel (functional.py:1452-1538) with the band of unclamped taps only.
inline bool build_resample_taps(int sr_in, ResamplePlan* plan, std::vector<int>* first, std::vector<float>* w) {
    long long a = sr_in, b = kIngestRate;
    while (b) { const long long t = a % b; a = b; b = t; }
    const int orig = static_cast<int>(sr_in / a), nw = static_cast<int>(kIngestRate / a);
    const double base = std::fmin(orig, nw) * 0.99;
    const int width = static_cast<int>(std::ceil(6.0 * orig / base));
    const int full = 2 * width + orig;
    const double scale = base / orig;
    std::vector<double> row(full);
    std::vector<int> lo(nw), hi(nw);
    int max_taps = 1;
    for (int pass = 0; pass < 2; ++pass) {
        if (pass == 1) {
            if (nw > 16384 || max_taps > 512) return false;
            first->assign(nw, 0);
            w->assign(static_cast<size_t>(nw) * max_taps, 0.f);
        }
        for (int p = 0; p < nw; ++p) {
            // the phase term is an int64 tensor divided by an int: float32 in torch, then promoted to float64
            const double phase = static_cast<double>(static_cast<float>(-p) / static_cast<float>(nw));
            int f = full, l = -1;
            for (int k = 0; k < full; ++k) {
                double t = (phase + static_cast<double>(k - width) / orig) * base;
                const bool clamped = t <= -6.0 || t >= 6.0;
                t = std::fmin(6.0, std::fmax(-6.0, t));
                const double win = std::cos(t * M_PI / 6.0 / 2.0);
                t *= M_PI;
                row[k] = (t == 0.0 ? 1.0 : std::sin(t) / t) * (win * win * scale);
                if (!clamped) {
                    if (k < f) f = k;
                    l = k;
                }
            }
            if (l < f) { f = 0; l = 0; }
            if (pass == 0) {
                if (l - f + 1 > max_taps) max_taps = l - f + 1;
            } else {
                if (f + max_taps > full) f = full - max_taps;           // keep the band inside the staged span
                if (f < 0) f = 0;
                (*first)[p] = f;
                for (int k = 0; k < max_taps && f + k < full; ++k)
                    (*w)[static_cast<size_t>(p) * max_taps + k] = static_cast<float>(row[f + k]);
            }
        }
    }
    plan->orig_f = orig;
    plan->new_f = nw;
    plan->width = width;
    plan->taps_full = full;
    plan->max_taps = max_taps;
    plan->first0 = (*first)[0];
    plan->first_spread = 0;
    for (int p = 0; p < nw; ++p)
        if ((*first)[p] - plan->first0 > plan->first_spread) plan->first_spread = (*first)[p] - plan->first0;
    auto start = [&](long long j) { return (j / nw) * orig + (*first)[j % nw]; };
    plan->pair_shift_max = 0;
    for (long long j = 0; j < 2LL * nw + 2; ++j) {
        if (start(j + 1) < start(j)) { plan->pair_shift_max = -1; break; }
        if (j % 2 == 0 && start(j + 1) - start(j) > plan->pair_shift_max) plan->pair_shift_max = static_cast<int>(start(j + 1) - start(j));
    }
    return true;
}

}  // namespace sad
