// 2048-point complex FFT on a GROUP of 128 threads, radix 16 x 16 x 8, fp32 -- the front end's transform (frontend.cu,
// round 2).  Each thread keeps 16 complex points in registers per pass, so the data crosses shared memory twice
// (after pass 1 and after pass 2) instead of three to four times with the radix 8-8-8-4 / 256-thread version
// (fft2048.cuh: 25% of the shared-memory wavefronts of that kernel were bank conflicts and the pipe was 73% busy).
//
// Two real frames are transformed at once (frame a in the real part, frame b in the imaginary part).
//
// Decimation-in-time Stockham passes, N = 2048, Ns = product of the radices already applied:
//   pass with radix R:  j in [0, N/R), k = j mod Ns,
//       v[q] = in[j + q*N/R] * W_{Ns*R}^{k*q},  V = DFT_R(v),  out[(j/Ns)*Ns*R + k + q*Ns] = V[q]
//   pass 1  R=16 Ns=1   thread j=t            reads x[t + 128q]            writes logical e1 = 16t + q
//   pass 2  R=16 Ns=16  thread j=t, k=t&15    reads e1 = t + 128q          writes logical e2 = (t>>4)*256 + k + 16q
//   pass 3  R=8  Ns=256 j in {t, t+128}       reads e2 = j + 256q          writes X[j + 256q]
// Storage: ONE buffer of complex slots, slot index s stored at pad(s) = s + (s >> 4) (8-byte elements: a half-warp
// touches 16 distinct bank pairs in every access below -- replayed on the CPU by fft16_host_check.cpp).
//   after pass 1: element e1 at slot e1.
//   pass 2 is IN PLACE: thread t reads slots t + 128q and writes V[q] back to slot t + 128q (no barrier between its loads
//     and stores); element e2 = (t>>4)*256 + (t&15) + 16q therefore sits at slot (e2 & 15) + 16*(e2 >> 8) + 128*((e2 >> 4) & 15).
//   pass 3 is IN PLACE too: X[j + 256q] sits at the slot its input e2 = j + 256q came from, i.e.
//     slot_of_bin(k) = (k & 15) + 16*(k >> 8) + 128*((k >> 4) & 15).
// Twiddles are per-thread CONSTANTS (k = t & 15 in pass 2, j in {t, t+128} in pass 3): 15 + 14 complex values held in
// registers for the whole kernel, exact table look-ups of exp(-2*pi*i*n/2048) -- never produced by repeated multiplication.
#pragma once
#include <cstdint>

#include "fft2048.cuh"   // cpx, cadd / csub / cmul / mul_neg_i, dft4, dft8, split-power helpers

namespace sad {

constexpr int kFft16Threads = 128;
constexpr int kFft16Slots = 2048 + 128;                    // pad(2047) = 2174

SAD_HD int pad16(int s) { return s + (s >> 4); }
SAD_HD int slot_of_bin(int k) { return (k & 15) + ((k >> 8) << 4) + (((k >> 4) & 15) << 7); }

// In-place 16-point DFT.  OUTPUT ORDER: X[ka + 4*kb] is left in v[kb + 4*ka] (use dft16_out(q) to index it).
SAD_HD int dft16_out(int q) { return (q >> 2) + ((q & 3) << 2); }
SAD_HD void dft16(cpx* v) {
    // n = 4*n1 + n0, k = ka + 4*kb:  X[k] = sum_n0 W16^(n0*ka) W4^(n0*kb) * ( sum_n1 x[4 n1 + n0] W4^(n1*ka) )
    const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f, h = 0.70710678118654752440f;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int n0 = 0; n0 < 4; ++n0) dft4(v[n0], v[n0 + 4], v[n0 + 8], v[n0 + 12]);   // v[n0 + 4 ka] = Y[n0][ka]
    // Y[n0][ka] *= W16^(n0*ka), W16 = exp(-2 pi i / 16)
    v[1 + 4] = cmul(v[1 + 4], cpx{c1, -s1});                   // W^1
    v[2 + 4] = cpx{h * (v[2 + 4].x + v[2 + 4].y), h * (v[2 + 4].y - v[2 + 4].x)};   // W^2 = (1 - i)/sqrt2
    v[3 + 4] = cmul(v[3 + 4], cpx{s1, -c1});                   // W^3
    v[1 + 8] = cpx{h * (v[1 + 8].x + v[1 + 8].y), h * (v[1 + 8].y - v[1 + 8].x)};   // W^2
    v[2 + 8] = mul_neg_i(v[2 + 8]);                            // W^4 = -i
    v[3 + 8] = cpx{h * (v[3 + 8].y - v[3 + 8].x), -h * (v[3 + 8].x + v[3 + 8].y)};  // W^6 = (-1 - i)/sqrt2
    v[1 + 12] = cmul(v[1 + 12], cpx{s1, -c1});                 // W^3
    v[2 + 12] = cpx{h * (v[2 + 12].y - v[2 + 12].x), -h * (v[2 + 12].x + v[2 + 12].y)};   // W^6
    v[3 + 12] = cmul(v[3 + 12], cpx{-c1, s1});                 // W^9
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int ka = 0; ka < 4; ++ka) dft4(v[4 * ka], v[4 * ka + 1], v[4 * ka + 2], v[4 * ka + 3]);   // v[kb + 4 ka] = X[ka + 4 kb]
}

// Twiddle angles (units of 2*pi/2048) of thread t: pass 2 q = 1..15 -> (t & 15) * q * 8; pass 3 butterfly h, q = 1..7 -> (t + 128h) * q.
SAD_HD int fft16_tw2_angle(int t, int q) { return ((t & 15) * q) << 3; }
SAD_HD int fft16_tw3_angle(int t, int h, int q) { return ((t + 128 * h) * q) & 2047; }

// ---- pass bodies of thread t (0..127); buf: complex slots (padded), tw2[15], tw3[2][7]: this thread's twiddles ----
SAD_HD void fft16_pass1(int t, cpx* v /*[16]: x[t + 128q], overwritten*/, cpx* buf) {
    dft16(v);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int q = 0; q < 16; ++q) buf[pad16(16 * t + q)] = v[dft16_out(q)];
}
template <typename Tw>
SAD_HD void fft16_pass2_tw(int t, Tw tw2 /* tw2(q) -> W_256^{(t & 15) q}, q = 1..15 */, cpx* buf) {
    cpx v[16];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int q = 0; q < 16; ++q) {
        const cpx x = buf[pad16(t + 128 * q)];
        v[q] = q == 0 ? x : cmul(x, tw2(q ? q : 1));
    }
    dft16(v);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int q = 0; q < 16; ++q) buf[pad16(t + 128 * q)] = v[dft16_out(q)];      // in place
}
SAD_HD void fft16_pass2(int t, const cpx* tw2, cpx* buf) {
    cpx v[16];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int q = 0; q < 16; ++q) {
        const cpx x = buf[pad16(t + 128 * q)];
        v[q] = q == 0 ? x : cmul(x, tw2[q ? q - 1 : 0]);
    }
    dft16(v);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int q = 0; q < 16; ++q) buf[pad16(t + 128 * q)] = v[dft16_out(q)];      // in place: V[q] -> the slot input q came from
}
template <typename Tw>
SAD_HD void fft16_pass3_tw(int t, Tw tw3 /* tw3(h, q) -> W_2048^{(t + 128 h) q}, q = 1..7 */, cpx* buf) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int h = 0; h < 2; ++h) {
        const int j = t + 128 * h;
        const int base = (j & 15) + ((j >> 4) << 7);               // slot of e2 = j + 256 q is base + 16 q
        cpx v[8];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int q = 0; q < 8; ++q) {
            const cpx x = buf[pad16(base + 16 * q)];
            v[q] = q == 0 ? x : cmul(x, tw3(h, q ? q : 1));
        }
        dft8(v);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int q = 0; q < 8; ++q)
            if (q != 4) buf[pad16(base + 16 * q)] = v[q];          // bins 1024..1279 (q = 4) are never read
    }
}
SAD_HD void fft16_pass3(int t, const cpx (*tw3)[7], cpx* buf) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int h = 0; h < 2; ++h) {
        const int j = t + 128 * h;
        const int base = (j & 15) + ((j >> 4) << 7);               // slot of e2 = j + 256 q is base + 16 q
        cpx v[8];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int q = 0; q < 8; ++q) {
            const cpx x = buf[pad16(base + 16 * q)];
            v[q] = q == 0 ? x : cmul(x, tw3[h][q ? q - 1 : 0]);
        }
        dft8(v);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int q = 0; q < 8; ++q)
            if (q != 4) buf[pad16(base + 16 * q)] = v[q];          // bins 1024..1279 (q = 4) are never read: the mel bank ends at bin 768
    }
}

// Power spectra of the two packed real frames at bin k (0..768) from the permuted spectrum in `buf`.
SAD_HD void fft16_split_power(const cpx* buf, int k, float& pa, float& pb) {
    const cpx z = buf[pad16(slot_of_bin(k))];
    const cpx w = buf[pad16(slot_of_bin((2048 - k) & 2047))];
    const float ar = 0.5f * (z.x + w.x), ai = 0.5f * (z.y - w.y);
    const float br = 0.5f * (z.y + w.y), bi = 0.5f * (w.x - z.x);
    pa = ar * ar + ai * ai;
    pb = br * br + bi * bi;
}

}  // namespace sad
