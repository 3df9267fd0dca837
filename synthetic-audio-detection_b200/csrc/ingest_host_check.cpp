// CPU check of the ingest stage's host arithmetic (there is no GPU in the build container):
//   ingest_host_check taps <sr_in>      -> "orig new width full max_taps", then per phase: first tap index and the taps
//   ingest_host_check length <sr_in> <n_frames>...  -> output length and un-padded length per n_frames
// tests/test_ingest_host.py compares both with the oracle's restatement of torchaudio (bit-exact float32 taps).
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "ingest_taps.h"

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
    fprintf(stderr, "usage: ingest_host_check taps <sr> | length <sr> <frames>...\n");
    return 2;
}
