// 2048-point complex FFT for a 256-thread CTA, Stockham auto-sort, radix 8-8-8-4, fp32.
//
// Used by the front-end kernel (frontend.cu) to transform TWO real 2048-sample frames at once (frame a in the
// real part, frame b in the imaginary part).  The per-thread pass bodies are __host__ __device__ so that the
// indexing, twiddles and shared-memory padding can be exercised on the CPU (tests/test_fft_host.py builds
// csrc/fft_host_check.cpp) -- there is no GPU in the build container.
//
// Pass structure for N = 2048 = 8*8*8*4, Ns = product of the radices already applied:
//   pass r with radix R:   j in [0, N/R):  k = j mod Ns
//       v[q]  = in[j + q*N/R] * exp(-2*pi*i * k*q / (Ns*R)),  q = 0..R-1
//       V     = DFT_R(v)
//       out[(j/Ns)*Ns*R + k + q*Ns] = V[q]
// Twiddles are table look-ups of exp(-2*pi*i*n/2048) (k*q < Ns*R always), never produced by repeated
// multiplication.  The table is stored PER PASS as tw[q-1][k] so that the lanes of a warp (consecutive k) read
// consecutive entries: indexing one W[2048] table by (k*q)*stride put every lane on the same bank (ncu: 54% of the
// kernel's shared-memory wavefronts were bank conflicts).
//
// Shared-memory layout: separate re/im float arrays, logical index a stored at pad(a):
//   after pass 1: pad1(a) = a + a/32        (stride-8 writes -> conflict free)
//   after pass 2: pad2(a) = a + 8*(a/64)    (writes in 4 groups of 8 spaced 64 -> conflict free)
//   after pass 3 and 4: no padding (writes are contiguous per warp)
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define SAD_HD __host__ __device__ __forceinline__
#else
#define SAD_HD inline
#endif

namespace sad {

constexpr int kFftN = 2048;
constexpr int kFftThreads = 256;
constexpr int kFftBuf = 2304;   // floats per component: max padded index is pad2(2047) = 2295

struct alignas(8) cpx {   // 8-byte aligned: one LDS.64 / STS.64 per element
    float x, y;
};

// Per-pass twiddle tables (entry n of the flattened struct has angle fft_twiddle_angle(n)): W_L^{k*q} with L = Ns*R of the pass.
struct FftTwiddles {
    cpx p2[7][8];      // pass 2: L = 64,   k = t & 7,  q = 1..7
    cpx p3[7][64];     // pass 3: L = 512,  k = t & 63, q = 1..7
    cpx p4[3][512];    // pass 4: L = 2048, k = j,      q = 1..3
};
constexpr int kFftTwiddleCount = 7 * 8 + 7 * 64 + 3 * 512;   // 2040 complex values


SAD_HD cpx cadd(cpx a, cpx b) { return {a.x + b.x, a.y + b.y}; }
SAD_HD cpx csub(cpx a, cpx b) { return {a.x - b.x, a.y - b.y}; }
SAD_HD cpx cmul(cpx a, cpx b) { return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
SAD_HD cpx mul_neg_i(cpx a) { return {a.y, -a.x}; }   // a * (-i)

SAD_HD int pad1(int a) { return a + (a >> 5); }
SAD_HD int pad2(int a) { return a + ((a >> 6) << 3); }

// In-place 4-point DFT, natural-order output.
SAD_HD void dft4(cpx& a0, cpx& a1, cpx& a2, cpx& a3) {
    cpx c0 = cadd(a0, a2), c1 = cadd(a1, a3);
    cpx d0 = csub(a0, a2), d1 = mul_neg_i(csub(a1, a3));
    a0 = cadd(c0, c1);
    a2 = csub(c0, c1);
    a1 = cadd(d0, d1);
    a3 = csub(d0, d1);
}

// In-place 8-point DFT (decimation in frequency), natural-order output.
SAD_HD void dft8(cpx* v) {
    const float h = 0.70710678118654752440f;
    cpx a0 = cadd(v[0], v[4]), a1 = cadd(v[1], v[5]), a2 = cadd(v[2], v[6]), a3 = cadd(v[3], v[7]);
    cpx b0 = csub(v[0], v[4]);
    cpx t1 = csub(v[1], v[5]);
    cpx t2 = csub(v[2], v[6]);
    cpx t3 = csub(v[3], v[7]);
    cpx b1 = {h * (t1.x + t1.y), h * (t1.y - t1.x)};    // * exp(-i*pi/4)  = (1 - i)/sqrt2
    cpx b2 = mul_neg_i(t2);                              // * exp(-i*pi/2)
    cpx b3 = {h * (t3.y - t3.x), -h * (t3.x + t3.y)};   // * exp(-3i*pi/4) = (-1 - i)/sqrt2
    dft4(a0, a1, a2, a3);   // even outputs X[0],X[2],X[4],X[6]
    dft4(b0, b1, b2, b3);   // odd outputs  X[1],X[3],X[5],X[7]
    v[0] = a0; v[2] = a1; v[4] = a2; v[6] = a3;
    v[1] = b0; v[3] = b1; v[5] = b2; v[7] = b3;
}

// ---- the four passes, body of thread t (0..255).  re/im: padded shared buffers; tw: W[2048] as (cos, -sin).
// Pass 1 takes its inputs from registers (the windowed samples t + 256*q of the two frames).
SAD_HD void fft_pass1(int t, const cpx* in8, float* re, float* im) {
    cpx v[8];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int q = 0; q < 8; ++q) v[q] = in8[q];
    dft8(v);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int q = 0; q < 8; ++q) {
        const int a = pad1(t * 8 + q);
        re[a] = v[q].x;
        im[a] = v[q].y;
    }
}
// Passes 2 and 3 are split into a load phase and a store phase because the transform is done in place:
// every thread must have read its inputs before any thread overwrites them (one barrier in between).
// n-th entry of the flattened FftTwiddles -> (pass table, q, k) -> angle index m of exp(-2*pi*i*m/2048)
SAD_HD int fft_twiddle_angle(int n) {
    if (n < 56) return ((n & 7) * (n / 8 + 1)) << 5;
    n -= 56;
    if (n < 448) return ((n & 63) * (n / 64 + 1)) << 2;
    n -= 448;
    return (n & 511) * (n / 512 + 1);
}

SAD_HD void fft_pass2_load(int t, const float* re, const float* im, const FftTwiddles& tw, cpx* v) {
    const int k = t & 7;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int q = 0; q < 8; ++q) {
        const int a = pad1(t + 256 * q);
        cpx x = {re[a], im[a]};
        v[q] = q == 0 ? x : cmul(x, tw.p2[q ? q - 1 : 0][k]);   // exp(-2 pi i k q / 64)
    }
    dft8(v);
}
SAD_HD void fft_pass2_store(int t, const cpx* v, float* re, float* im) {
    const int base = ((t >> 3) << 6) + (t & 7);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int q = 0; q < 8; ++q) {
        const int a = pad2(base + 8 * q);
        re[a] = v[q].x;
        im[a] = v[q].y;
    }
}
SAD_HD void fft_pass3_load(int t, const float* re, const float* im, const FftTwiddles& tw, cpx* v) {
    const int k = t & 63;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int q = 0; q < 8; ++q) {
        const int a = pad2(t + 256 * q);
        cpx x = {re[a], im[a]};
        v[q] = q == 0 ? x : cmul(x, tw.p3[q ? q - 1 : 0][k]);   // exp(-2 pi i k q / 512)
    }
    dft8(v);
}
SAD_HD void fft_pass3_store(int t, const cpx* v, float* re, float* im) {
    const int base = ((t >> 6) << 9) + (t & 63);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int q = 0; q < 8; ++q) {
        re[base + 64 * q] = v[q].x;
        im[base + 64 * q] = v[q].y;
    }
}
// Pass 4: radix 4, Ns = 512; thread t handles j = t and j = t + 256.  Output is the natural-order spectrum.
SAD_HD void fft_pass4_load(int t, const float* re, const float* im, const FftTwiddles& tw, cpx* v /*[8]*/) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int h = 0; h < 2; ++h) {
        const int j = t + 256 * h;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int q = 0; q < 4; ++q) {
            cpx x = {re[j + 512 * q], im[j + 512 * q]};
            v[4 * h + q] = q == 0 ? x : cmul(x, tw.p4[q ? q - 1 : 0][j]);   // exp(-2 pi i j q / 2048)
        }
        dft4(v[4 * h], v[4 * h + 1], v[4 * h + 2], v[4 * h + 3]);
    }
}
SAD_HD void fft_pass4_store(int t, const cpx* v, float* re, float* im) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int h = 0; h < 2; ++h) {
        const int j = t + 256 * h;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int q = 0; q < 4; ++q) {
            re[j + 512 * q] = v[4 * h + q].x;
            im[j + 512 * q] = v[4 * h + q].y;
        }
    }
}

// Power spectra of the two real frames packed as Z = FFT(a + i b):
//   A[k] = (Z[k] + conj(Z[N-k]))/2,  B[k] = (Z[k] - conj(Z[N-k]))/(2i)
SAD_HD void split_power(const float* re, const float* im, int k, float& pa, float& pb) {
    const int kn = (kFftN - k) & (kFftN - 1);
    const float zr = re[k], zi = im[k], wr = re[kn], wi = im[kn];
    const float ar = 0.5f * (zr + wr), ai = 0.5f * (zi - wi);
    const float br = 0.5f * (zi + wi), bi = 0.5f * (wr - zr);
    pa = ar * ar + ai * ai;
    pb = br * br + bi * bi;
}

}  // namespace sad
