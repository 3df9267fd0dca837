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

// ---- few-phase ratios: the taps as kernel parameters -----------------------------------------------------------------------
// With one or two phases (96 / 64 / 192 kHz: one; 48 / 16 kHz: two) every thread of a block filters the same phases, so
// the taps need no registers: they travel as a kernel parameter and reach the FMAs as constant-bank operands.  A thread
// then affords `outputs` adjacent outputs on one window (eight at 48 kHz: one 16-byte shared load per output instead of
// three).  w[p][taps] holds phase p shifted to its place in that window: w[p][lead + first[p] - first0 + k] = tap k.
struct UniformTaps {
    int orig, phases, outputs, taps;   // 0 outputs: not a few-phase ratio
    float w[80];                       // [phases][taps]
};

struct UniformConfig { int orig, phases, outputs, taps; };
// outputs: a multiple of phases with (outputs / phases) * orig % 4 == 0, so that every thread's window starts on a
// 16-byte boundary; taps: max_taps + the largest shift, rounded up to 4
constexpr UniformConfig kUniformConfigs[] = {{3, 1, 4, 40}, {2, 1, 4, 28}, {6, 1, 2, 76}, {3, 2, 8, 24}, {1, 2, 8, 16}};

inline int even_lead(const ResamplePlan& plan) { return ((plan.first0 - plan.width) % 2 + 2) % 2; }

inline bool build_uniform_taps(const ResamplePlan& plan, const std::vector<int>& first, const std::vector<float>& w, UniformTaps* u) {
    *u = UniformTaps{};
    if (plan.pair_shift_max < 0) return false;
    for (const UniformConfig& cfg : kUniformConfigs) {
        if (cfg.orig != plan.orig_f || cfg.phases != plan.new_f) continue;
        const int lead = even_lead(plan);
        for (int p = 0; p < cfg.phases; ++p) {
            const int shift = lead + first[p] - plan.first0;
            if (shift < 0 || shift + plan.max_taps > cfg.taps) return false;
            for (int k = 0; k < plan.max_taps; ++k) u->w[p * cfg.taps + shift + k] = w[static_cast<size_t>(p) * plan.max_taps + k];
        }
        u->orig = cfg.orig;
        u->phases = cfg.phases;
        u->outputs = cfg.outputs;
        u->taps = cfg.taps;
        return true;
    }
    return false;
}

// ---- geometry of the staged (`pair`) resampling kernel of ingest.cu, chosen on the host -----------------------------------
// Kept here, free of CUDA, so that csrc/ingest_host_check.cpp can replay the kernel's index walk on the CPU.
constexpr int kPairDepth = 2;        // slots of the raw-PCM ring in shared memory
constexpr int kPairThreads = 320;    // largest block

struct PairGeometry {
    int rounds;          // rounds of G * blockDim outputs per item
    int round_stride;    // input frames between the sub-spans of two rounds
    int sub_floats;      // floats per sub-span (a multiple of 4): round_stride + overlap, at least
    int overlap;         // frames of the next round a sub-span also holds
    int n_chunks;        // 16-byte chunks of raw PCM per item
    int two;             // 1: round_stride is even, the conversion pass handles two frames per step (8-byte accesses)
    int lead;            // 0 / 1 extra frame staged in front of every item so that its first frame has an even index
};

struct PairChoice {
    int uniform;         // 1: the few-phase kernel (taps as kernel parameters, see UniformTaps)
    int outputs;         // G: adjacent outputs per thread (2, or 1 for the 34-37-tap rates; UniformTaps::outputs)
    int window;          // TE: floats of the per-thread window (20, 24, 28 or 40; few-phase: (outputs / phases - 1) * orig + taps)
    int threads;         // block size: whole periods of the phase pattern
    PairGeometry geo;
    size_t smem;         // dynamic shared memory per block
};

// False when the ratio does not fit the kernel (band starts not monotone, window > 40 floats, phase period > 320 threads).
// bytes_per_frame: 2, 4 or 8 (mono / stereo, int16 / float32).
inline bool choose_pair_geometry(const ResamplePlan& plan, int bytes_per_frame, long long out_len, int sms, PairChoice* c,
                                 const UniformTaps* uni = nullptr) {
    if (plan.pair_shift_max < 0) return false;
    const bool uniform = uni && uni->outputs > 0;
    int G = 2, need = 3 + plan.pair_shift_max + plan.max_taps;     // alignment slack of a 16-byte window + the pair's shift + taps
    int TE, threads;
    if (uniform) {
        G = uni->outputs;
        TE = ((G / uni->phases - 1) * uni->orig + uni->taps + 3) / 4 * 4;
        threads = 256;
    } else {
        if (need > 28) {
            G = 1;
            need = 3 + plan.max_taps;
            if (need > 40) return false;
        }
        TE = G == 1 ? 40 : need <= 20 ? 20 : need <= 24 ? 24 : 28;
        const int period = G == 2 && plan.new_f % 2 == 0 ? plan.new_f / 2 : plan.new_f;   // threads per period of the phase pattern
        if (period > kPairThreads) return false;
        threads = kPairThreads / period * period;
        if (threads < 128) return false;
    }
    const int fpc = 16 / bytes_per_frame;
    PairGeometry geo{};
    size_t smem = 0;
    int rounds = 0;
    for (int attempt = 0; attempt < 2 && rounds < 1; ++attempt) {
        if (attempt == 1) {
            if (!uniform) return false;
            threads = 128;                                           // float stereo: half the block, still four per SM
        }
        const int frames_round = G * threads / plan.new_f;
        geo = PairGeometry{};
        geo.round_stride = frames_round * plan.orig_f;
        // items start round_stride * rounds frames apart: with an even stride the parity of an item's first frame is that
        // of first0 - width for every item, and one leading frame makes it even
        geo.two = geo.round_stride % 2 == 0;
        geo.lead = geo.two ? even_lead(plan) : 0;
        if (uniform && !geo.two) return false;                       // (cannot happen: the thread stride is a multiple of 4)
        // frames a round's windows reach over; few-phase: thread t's window is [t * stride, + TE)
        const int need_round = uniform ? (threads - 1) * (G / plan.new_f) * plan.orig_f + TE
                                       : geo.lead + (frames_round - 1) * plan.orig_f + plan.first_spread + TE;
        geo.sub_floats = (need_round + 3) / 4 * 4;
        if (geo.sub_floats < geo.round_stride) geo.sub_floats = (geo.round_stride + 3) / 4 * 4;
        geo.overlap = geo.sub_floats - geo.round_stride;
        auto size_for = [&](int r) {
            geo.rounds = r;
            const long long need_item = static_cast<long long>(r - 1) * geo.round_stride + geo.sub_floats;
            geo.n_chunks = static_cast<int>((need_item + fpc - 1 + fpc - 1) / fpc);   // the first chunk may start fpc - 1 frames early
            smem = static_cast<size_t>(r) * geo.sub_floats * 4 + static_cast<size_t>(kPairDepth) * geo.n_chunks * 16;
            return smem;
        };
        const size_t budget = (uniform ? 52 : 100) * 1024;           // two blocks per SM (four of the few-phase kernel)
        rounds = 16;
        while (rounds >= 1 && size_for(rounds) > budget) rounds >>= 1;
        // short streams: smaller items, so that every resident block gets a few
        const long long per_round = static_cast<long long>(G) * threads;
        while (rounds > 1 && (out_len + per_round * rounds - 1) / (per_round * rounds) < 8LL * sms) rounds >>= 1;
        if (rounds >= 1) size_for(rounds);
    }
    if (rounds < 1) return uniform ? choose_pair_geometry(plan, bytes_per_frame, out_len, sms, c) : false;
    c->uniform = uniform ? 1 : 0;
    c->outputs = G;
    c->window = TE;
    c->threads = threads;
    c->geo = geo;
    c->smem = smem;
    return true;
}

}  // namespace sad
