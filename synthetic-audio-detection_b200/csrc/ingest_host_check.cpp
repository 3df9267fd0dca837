// CPU check of the ingest stage's host arithmetic (there is no GPU in the build container):
//   ingest_host_check taps <sr_in>      -> "orig new width full max_taps", then per phase: first tap index and the taps
//   ingest_host_check length <sr_in> <n_frames>...  -> output length and un-padded length per n_frames
//   ingest_host_check pair <sr_in> <bytes_per_frame> <n_frames> <sms> [few_phase = 1]
//       -> replays the index walk of the staged (`pair`) resampling kernel of ingest.cu with the geometry the host picks
//          (choose_pair_geometry): raw chunks of an item -> per-round float sub-spans -> per-thread windows, and checks
//          that tap k of every output reads exactly the frame torchaudio's kernel reads; prints the geometry and "ok <n>".
// tests/test_ingest_host.py compares the first two with the oracle's restatement of torchaudio (bit-exact float32 taps)
// and runs the third over the rate / format / length grid.
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>

#include "ingest_taps.h"

// Frames are identified by their index in the stream; kOutside marks frames the kernel reads as zero (before the
// first / after the last) and kUnset marks shared-memory floats nothing was written to.
static const long long kOutside = -1, kUnset = -2;

static int skip_of(long long ff, int fpc) { return static_cast<int>(ff & (fpc - 1)); }

static int replay_pair(int sr, int bytes_per_frame, long long n_frames, int sms, bool few_phase) {
    sad::ResamplePlan plan{};
    std::vector<int> first;
    std::vector<float> w;
    if (!sad::build_resample_taps(sr, &plan, &first, &w)) {
        printf("unsupported\n");
        return 0;
    }
    long long n_real = 0;
    const long long out_len = sad::ingest_length(n_frames, sr, &n_real);
    sad::PairChoice c{};
    sad::UniformTaps uni{};
    const bool has_uni = few_phase && sad::build_uniform_taps(plan, first, w, &uni);
    if (has_uni) {                                                                 // the parameter block holds exactly the shifted taps
        std::vector<float> want(80, 0.f);
        for (int p = 0; p < uni.phases; ++p)
            for (int k = 0; k < plan.max_taps; ++k)
                want[p * uni.taps + sad::even_lead(plan) + first[p] - plan.first0 + k] = w[static_cast<size_t>(p) * plan.max_taps + k];
        if (uni.phases * uni.taps > 80 || memcmp(want.data(), uni.w, sizeof(uni.w)) != 0 || uni.outputs % uni.phases ||
            (uni.outputs / uni.phases) * uni.orig % 4) {
            printf("bad uniform taps\n");
            return 1;
        }
    }
    if (!sad::choose_pair_geometry(plan, bytes_per_frame, out_len, sms, &c, has_uni ? &uni : nullptr)) {
        printf("not a pair ratio\n");
        return 0;
    }
    const sad::PairGeometry& g = c.geo;
    const int G = c.outputs, TE = c.window, nt = c.threads, fpc = 16 / bytes_per_frame;
    printf("G %d TE %d threads %d rounds %d stride %d sub %d overlap %d chunks %d smem %zu two %d lead %d uniform %d\n", G, TE, nt,
           g.rounds, g.round_stride, g.sub_floats, g.overlap, g.n_chunks, c.smem, g.two, g.lead, c.uniform);
    if (G * nt % plan.new_f || g.sub_floats % 4 || g.overlap != g.sub_floats - g.round_stride || c.smem > 113 * 1024) {
        printf("bad geometry\n");
        return 1;
    }
    const long long per_item = static_cast<long long>(G) * nt * g.rounds;
    const long long n_items = (out_len + per_item - 1) / per_item;
    const long long frames_item = static_cast<long long>(g.round_stride) * g.rounds;
    std::vector<long long> raw(static_cast<size_t>(g.n_chunks) * fpc), span(static_cast<size_t>(g.rounds) * g.sub_floats);
    long long checked = 0;
    for (long long item = 0; item < n_items; ++item) {
        const long long ff = item * frames_item + plan.first0 - plan.width - g.lead;   // first frame staged for the item
        if (g.two && ((ff & 1) || (g.round_stride & 1) || (skip_of(ff, fpc) & 1))) {
            printf("two-frame conversion on an odd frame (item %lld)\n", item);
            return 1;
        }
        const long long a_lo = ff & ~static_cast<long long>(fpc - 1);
        const int skip = static_cast<int>(ff & (fpc - 1));
        for (int i = 0; i < g.n_chunks; ++i)                                       // issue(): 16-byte chunks, zero-filled outside
            for (int e = 0; e < fpc; ++e) {
                const long long f = a_lo + static_cast<long long>(fpc) * i + e;
                raw[static_cast<size_t>(i) * fpc + e] = (f >= 0 && f < n_frames) ? f : kOutside;
            }
        std::fill(span.begin(), span.end(), kUnset);
        for (int k = 0; k < g.sub_floats; ++k)                                     // conversion pass
            for (int i = 0; i < g.rounds; ++i) {
                const size_t src = static_cast<size_t>(skip) + static_cast<size_t>(i) * g.round_stride + k;
                if (src >= raw.size()) {
                    printf("conversion reads past the staged chunks (item %lld round %d k %d)\n", item, i, k);
                    return 1;
                }
                span[static_cast<size_t>(i) * g.sub_floats + k] = raw[src];
            }
        for (int t = 0; t < nt; ++t) {                                             // filter pass
            int pos[16], ph[16];
            int base;
            if (c.uniform) {
                base = t * (G / uni.phases) * uni.orig;
                if (base % 4) {
                    printf("window of thread %d is not 16-byte aligned\n", t);
                    return 1;
                }
                for (int q = 0; q < G; ++q) {
                    ph[q] = q % uni.phases;
                    pos[q] = base + (q / uni.phases) * uni.orig + g.lead + first[ph[q]] - plan.first0;   // where w[ph][shift + k] meets x
                    if ((G * t + q) % plan.new_f != ph[q]) {
                        printf("phase of thread %d output %d\n", t, q);
                        return 1;
                    }
                }
            } else {
                for (int q = 0; q < G; ++q) {
                    const int m = (G * t + q) / plan.new_f;
                    ph[q] = (G * t + q) - m * plan.new_f;
                    pos[q] = g.lead + m * plan.orig_f + first[ph[q]] - plan.first0;
                }
                base = pos[0] & ~3;
            }
            if (base + TE > g.sub_floats) {
                printf("window of thread %d leaves its sub-span\n", t);
                return 1;
            }
            for (int q = 0; q < G; ++q) {
                const int sh = pos[q] - base;
                if (sh < 0 || sh + plan.max_taps > TE) {
                    printf("taps of thread %d output %d do not fit the window (shift %d)\n", t, q, sh);
                    return 1;
                }
                for (int i = 0; i < g.rounds; ++i) {
                    const long long j = item * per_item + static_cast<long long>(G) * nt * i + G * t + q;
                    if (j >= n_real) continue;                                      // written as zero padding
                    const long long m = j / plan.new_f;
                    for (int k = 0; k < plan.max_taps; ++k) {
                        const long long want = m * plan.orig_f + first[ph[q]] + k - plan.width;   // torchaudio's xpad index - width
                        const long long got = span[static_cast<size_t>(i) * g.sub_floats + base + sh + k];
                        const long long expect = (want >= 0 && want < n_frames) ? want : kOutside;
                        if (got != expect) {
                            printf("output %lld tap %d reads frame %lld, wants %lld\n", j, k, got, expect);
                            return 1;
                        }
                    }
                    ++checked;
                }
            }
        }
    }
    if (checked != n_real) {
        printf("covered %lld of %lld outputs\n", checked, n_real);
        return 1;
    }
    printf("ok %lld\n", checked);
    return 0;
}

int main(int argc, char** argv) {
    if (argc >= 3 && !strcmp(argv[1], "taps")) {
        sad::ResamplePlan plan{};
        std::vector<int> first;
        std::vector<float> w;
        if (!sad::build_resample_taps(atoi(argv[2]), &plan, &first, &w)) {
            printf("unsupported\n");
            return 0;
        }
        printf("%d %d %d %d %d\n", plan.orig_f, plan.new_f, plan.width, plan.taps_full, plan.max_taps);
        for (int p = 0; p < plan.new_f; ++p) {
            printf("%d", first[p]);
            for (int k = 0; k < plan.max_taps; ++k) {
                unsigned bits;
                memcpy(&bits, &w[static_cast<size_t>(p) * plan.max_taps + k], 4);
                printf(" %08x", bits);
            }
            printf("\n");
        }
        return 0;
    }
    if (argc >= 4 && !strcmp(argv[1], "length")) {
        const int sr = atoi(argv[2]);
        for (int i = 3; i < argc; ++i) {
            long long n_real = 0;
            const long long n = sad::ingest_length(atoll(argv[i]), sr, &n_real);
            printf("%lld %lld\n", n, n_real);
        }
        return 0;
    }
    if (argc >= 6 && !strcmp(argv[1], "pair"))
        return replay_pair(atoi(argv[2]), atoi(argv[3]), atoll(argv[4]), atoi(argv[5]), argc < 7 || atoi(argv[6]) != 0);
    fprintf(stderr, "usage: ingest_host_check taps <sr> | length <sr> <frames>... | pair <sr> <bytes_per_frame> <frames> <sms>\n");
    return 2;
}
