// CPU exercise of the FFT pass bodies in fft2048.cuh (test helper, not part of the product library).
// Emulates the 256 threads of a CTA pass by pass and compares against a float64 DFT of two real frames.
// Also replays the shared-memory bank mapping of every store/load to prove the padding is conflict free.
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "fft2048.cuh"

using namespace sad;

static int max_conflict(const std::vector<int>& addrs) {   // addrs: 32 word addresses of one warp access
    int cnt[32] = {0};
    int worst = 0;
    for (int a : addrs) worst = std::max(worst, ++cnt[a & 31]);
    return worst;
}

int main() {
    static FftTwiddles twt;
    {
        cpx* flat = reinterpret_cast<cpx*>(&twt);
        for (int n = 0; n < kFftTwiddleCount; ++n) {
            double ang = -2.0 * M_PI * fft_twiddle_angle(n) / kFftN;
            flat[n] = {(float)std::cos(ang), (float)std::sin(ang)};
        }
    }
    std::vector<float> a(kFftN), b(kFftN);
    srand(7);
    for (int n = 0; n < kFftN; ++n) {
        a[n] = (float)(rand() / (double)RAND_MAX - 0.5) + 0.4f * (float)std::sin(2 * M_PI * 123.25 * n / kFftN);
        b[n] = (float)(rand() / (double)RAND_MAX - 0.5) * 1e-3f + 0.5f * (float)std::cos(2 * M_PI * 700.5 * n / kFftN);
    }
    std::vector<float> re(kFftBuf, 0.f), im(kFftBuf, 0.f);
    std::vector<cpx> regs(kFftThreads * 8);
    for (int t = 0; t < kFftThreads; ++t) {
        cpx in8[8];
        for (int q = 0; q < 8; ++q) in8[q] = {a[t + 256 * q], b[t + 256 * q]};
        fft_pass1(t, in8, re.data(), im.data());
    }
    for (int t = 0; t < kFftThreads; ++t) fft_pass2_load(t, re.data(), im.data(), twt, &regs[t * 8]);
    for (int t = 0; t < kFftThreads; ++t) fft_pass2_store(t, &regs[t * 8], re.data(), im.data());
    for (int t = 0; t < kFftThreads; ++t) fft_pass3_load(t, re.data(), im.data(), twt, &regs[t * 8]);
    for (int t = 0; t < kFftThreads; ++t) fft_pass3_store(t, &regs[t * 8], re.data(), im.data());
    for (int t = 0; t < kFftThreads; ++t) fft_pass4_load(t, re.data(), im.data(), twt, &regs[t * 8]);
    for (int t = 0; t < kFftThreads; ++t) fft_pass4_store(t, &regs[t * 8], re.data(), im.data());

    // float64 reference
    double max_rel = 0, peak = 0;
    std::vector<double> pa_ref(1025), pb_ref(1025);
    for (int k = 0; k <= 1024; ++k) {
        std::complex<double> sa = 0, sb = 0;
        for (int n = 0; n < kFftN; ++n) {
            std::complex<double> w = std::polar(1.0, -2.0 * M_PI * ((long long)k * n % kFftN) / kFftN);
            sa += (double)a[n] * w;
            sb += (double)b[n] * w;
        }
        pa_ref[k] = std::norm(sa);
        pb_ref[k] = std::norm(sb);
        peak = std::max(peak, std::max(pa_ref[k], pb_ref[k]));
    }
    double worst_a = 0, worst_b = 0;
    for (int k = 0; k <= 1024; ++k) {
        float pa, pb;
        split_power(re.data(), im.data(), k, pa, pb);
        worst_a = std::max(worst_a, std::fabs(pa - pa_ref[k]) / (pa_ref[k] + 1e-7 * peak));
        worst_b = std::max(worst_b, std::fabs(pb - pb_ref[k]) / (pb_ref[k] + 1e-7 * peak));
    }
    max_rel = std::max(worst_a, worst_b);
    printf("max_rel_power_err %.3e (a %.3e, b %.3e)\n", max_rel, worst_a, worst_b);

    // bank-conflict replay (4-byte words, 32 banks), per warp and per q
    int worst = 1;
    for (int w = 0; w < 8; ++w)
        for (int q = 0; q < 8; ++q) {
            std::vector<int> s1, l2, s2, l3, s3;
            for (int l = 0; l < 32; ++l) {
                int t = w * 32 + l;
                s1.push_back(pad1(t * 8 + q));
                l2.push_back(pad1(t + 256 * q));
                s2.push_back(pad2(((t >> 3) << 6) + (t & 7) + 8 * q));
                l3.push_back(pad2(t + 256 * q));
                s3.push_back(((t >> 6) << 9) + (t & 63) + 64 * q);
            }
            worst = std::max(worst, std::max(max_conflict(s1), std::max(max_conflict(l2), std::max(max_conflict(s2),
                             std::max(max_conflict(l3), max_conflict(s3))))));
        }
    // twiddle reads are 8-byte: hardware serves a warp in two half-warp passes; within each, distinct 8-byte
    // addresses must fall on distinct bank PAIRS (identical addresses broadcast)
    auto conflict64 = [](const std::vector<long>& byte_addr) {
        int w = 1;
        for (int half = 0; half < 2; ++half) {
            int cnt[16] = {0};
            std::vector<long> seen;
            for (int l = 16 * half; l < 16 * half + 16; ++l) {
                bool dup = false;
                for (long a : seen) dup |= (a == byte_addr[l]);
                if (dup) continue;
                seen.push_back(byte_addr[l]);
                w = std::max(w, ++cnt[(byte_addr[l] / 8) & 15]);
            }
        }
        return w;
    };
    int worst_tw = 1;
    const char* base = reinterpret_cast<const char*>(&twt);
    for (int w = 0; w < 8; ++w)
        for (int q = 1; q < 8; ++q) {
            std::vector<long> a2, a3, a4a, a4b;
            for (int l = 0; l < 32; ++l) {
                int t = w * 32 + l;
                a2.push_back(reinterpret_cast<const char*>(&twt.p2[q - 1][t & 7]) - base);
                a3.push_back(reinterpret_cast<const char*>(&twt.p3[q - 1][t & 63]) - base);
                if (q < 4) {
                    a4a.push_back(reinterpret_cast<const char*>(&twt.p4[q - 1][t]) - base);
                    a4b.push_back(reinterpret_cast<const char*>(&twt.p4[q - 1][t + 256]) - base);
                }
            }
            worst_tw = std::max(worst_tw, std::max(conflict64(a2), conflict64(a3)));
            if (q < 4) worst_tw = std::max(worst_tw, std::max(conflict64(a4a), conflict64(a4b)));
        }
    printf("max_twiddle_conflict %d\n", worst_tw);
    worst = std::max(worst, worst_tw);
    printf("max_bank_conflict %d\n", worst);
    int max_index = 0;
    for (int a2 = 0; a2 < kFftN; ++a2) max_index = std::max(max_index, std::max(pad1(a2), pad2(a2)));
    printf("max_padded_index %d (buffer %d)\n", max_index, kFftBuf);
    return (max_rel < 5e-4 && worst == 1 && max_index < kFftBuf) ? 0 : 1;
}
