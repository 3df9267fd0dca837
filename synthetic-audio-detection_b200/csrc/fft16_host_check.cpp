// CPU exercise of the radix 16-16-8 FFT pass bodies in fft2048r16.cuh (test helper, not part of the product library).
// Emulates the 128 threads of a group pass by pass, compares the power spectra of two packed real frames against a
// float64 DFT, and replays the shared-memory bank mapping of every 8-byte access (per half-warp: 16 distinct bank pairs).
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "fft2048r16.cuh"

using namespace sad;

static int conflict64(const std::vector<int>& slot) {   // 32 padded slot indices (8-byte elements) of one warp access
    int worst = 1;
    for (int half = 0; half < 2; ++half) {
        int cnt[16] = {0};
        std::vector<int> seen;
        for (int l = 16 * half; l < 16 * half + 16; ++l) {
            if (std::find(seen.begin(), seen.end(), slot[l]) != seen.end()) continue;
            seen.push_back(slot[l]);
            worst = std::max(worst, ++cnt[slot[l] & 15]);
        }
    }
    return worst;
}

int main() {
    const int N = 2048, T = kFft16Threads;
    std::vector<float> a(N), b(N);
    srand(7);
    for (int n = 0; n < N; ++n) {
        a[n] = (float)(rand() / (double)RAND_MAX - 0.5) + 0.4f * (float)std::sin(2 * M_PI * 123.25 * n / N);
        b[n] = (float)(rand() / (double)RAND_MAX - 0.5) * 1e-3f + 0.5f * (float)std::cos(2 * M_PI * 700.5 * n / N);
    }
    auto tw = [](int angle) {
        const double x = -2.0 * M_PI * angle / 2048.0;
        return cpx{(float)std::cos(x), (float)std::sin(x)};
    };
    std::vector<cpx> buf(kFft16Slots, cpx{0.f, 0.f});
    for (int t = 0; t < T; ++t) {
        cpx v[16];
        for (int q = 0; q < 16; ++q) v[q] = {a[t + 128 * q], b[t + 128 * q]};
        fft16_pass1(t, v, buf.data());
    }
    for (int t = 0; t < T; ++t) {
        cpx tw2[15];
        for (int q = 1; q < 16; ++q) tw2[q - 1] = tw(fft16_tw2_angle(t, q));
        fft16_pass2(t, tw2, buf.data());
    }
    for (int t = 0; t < T; ++t) {
        cpx tw3[2][7];
        for (int h = 0; h < 2; ++h)
            for (int q = 1; q < 8; ++q) tw3[h][q - 1] = tw(fft16_tw3_angle(t, h, q));
        fft16_pass3(t, tw3, buf.data());
    }
    double peak = 0;
    std::vector<double> pa_ref(769), pb_ref(769);
    for (int k = 0; k <= 768; ++k) {
        std::complex<double> sa = 0, sb = 0;
        for (int n = 0; n < N; ++n) {
            std::complex<double> w = std::polar(1.0, -2.0 * M_PI * ((long long)k * n % N) / N);
            sa += (double)a[n] * w;
            sb += (double)b[n] * w;
        }
        pa_ref[k] = std::norm(sa);
        pb_ref[k] = std::norm(sb);
        peak = std::max(peak, std::max(pa_ref[k], pb_ref[k]));
    }
    double worst_a = 0, worst_b = 0;
    for (int k = 0; k <= 768; ++k) {
        float pa, pb;
        fft16_split_power(buf.data(), k, pa, pb);
        worst_a = std::max(worst_a, std::fabs(pa - pa_ref[k]) / (pa_ref[k] + 1e-7 * peak));
        worst_b = std::max(worst_b, std::fabs(pb - pb_ref[k]) / (pb_ref[k] + 1e-7 * peak));
    }
    printf("max_rel_power_err %.3e (a %.3e, b %.3e)\n", std::max(worst_a, worst_b), worst_a, worst_b);

    int worst = 1;
    for (int w = 0; w < 4; ++w) {
        for (int q = 0; q < 16; ++q) {
            std::vector<int> s1, a2, a3, a3b;
            for (int l = 0; l < 32; ++l) {
                const int t = w * 32 + l;
                s1.push_back(pad16(16 * t + q));
                a2.push_back(pad16(t + 128 * q));
                if (q < 8) {
                    a3.push_back(pad16((t & 15) + ((t >> 4) << 7) + 16 * q));
                    const int j = t + 128;
                    a3b.push_back(pad16((j & 15) + ((j >> 4) << 7) + 16 * q));
                }
            }
            worst = std::max(worst, std::max(conflict64(s1), conflict64(a2)));
            if (q < 8) worst = std::max(worst, std::max(conflict64(a3), conflict64(a3b)));
        }
        for (int i = 0; i < 7; ++i) {               // power stage: k = t + 128 i and its mirror
            std::vector<int> z, m;
            for (int l = 0; l < 32; ++l) {
                const int k = std::min(w * 32 + l + 128 * i, 768);
                z.push_back(pad16(slot_of_bin(k)));
                m.push_back(pad16(slot_of_bin((2048 - k) & 2047)));
            }
            worst = std::max(worst, std::max(conflict64(z), conflict64(m)));
        }
    }
    printf("max_bank_conflict %d\n", worst);
    int max_index = 0;
    for (int s = 0; s < N; ++s) max_index = std::max(max_index, pad16(s));
    printf("max_padded_index %d (buffer %d)\n", max_index, kFft16Slots);
    return (std::max(worst_a, worst_b) < 5e-4 && worst <= 2 && max_index < kFft16Slots) ? 0 : 1;
}
