// K2: fused stem for the single-channel log-mel image:
//     conv 7x7 / stride 2 / pad 3 (1 -> 64 channels per head, BN folded) + ReLU + maxpool 3x3 / stride 2 / pad 1
// for TWO heads at a time (N = 128), never materialising the 64 x 256 x 256 conv output (8 MB per head and segment).
// Replaces conv1 / bn1 / act1 / maxpool of timm's ResNet inside BinaryClassifier.forward (reference
// modular/source/inference_runner.py:49-51); the three identical input channels (:173) are folded into one by
// summing conv1's weights over Cin (api.cu).
//
// GEMM view per conv-output row r (256 pixels): two M=128 tiles, "even" (row i <-> conv pixel x = 2i) and "odd"
// (row i <-> x = 2i+1), so that TMEM lane i holds exactly the conv pixels pooled output px = i needs (2i, 2i+1 in its
// own lane, 2i-1 in lane i-1).  K = 64: chunk ky (16 B) of a row = the 8 image pixels [2x-3, 2x+4] of image row
// 2r+ky-3 (7 taps + one zero-weight slot); chunk 7 = {1,1,1,0,...} against {bias_hi, bias_mid, bias_lo} so the folded
// BN shift is added by the tensor core in fp32 (bias split into three bf16 terms, exact to 24 bits).
//   builders (4 warps): read the bf16 image (L2 resident) with 8-byte loads, funnel-shift, write both tiles into
//                       SWIZZLE_128B shared memory; double buffered.
//   MMA (1 thread)    : 8 x tcgen05.mma 128x128x16 per conv row into TMEM (even: cols 0-127, odd: 128-255; x2 buffers).
//   epilogue (4 warps): h = max(even, odd, odd of lane-1) per conv row (cross-warp lane via a tiny smem exchange),
//                       3-row vertical max carried in registers as packed bf16, ReLU, bf16, TMA store of the
//                       pooled row [32 px][64 ch] per warp and head.
// Work unit = (image, head pair, strip of 32 pooled rows); persistent CTAs, round-robin.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "ptx.cuh"
#include "stem_fused.h"

namespace sad {

namespace {

constexpr int kThreads = 288;                  // warp 0: MMA + TMA; warps 1-4: builders; warps 5-8: epilogue
constexpr int kStripRows = 32;                 // pooled rows per unit
constexpr int kStrips = 128 / kStripRows;
constexpr int kConvRowsPerUnit = 2 * kStripRows + 1;
constexpr int kTile = 128 * 128;               // one A tile (128 rows x 128 B) = one weight block (128 co x 64 k)
constexpr int kABytes = 2 /*bufs*/ * 2 /*even,odd*/ * kTile;     // 64 KB
constexpr int kWBytes = 2 * kTile;                                // 32 KB (double buffered)
constexpr int kOutBytes = 4 /*warps*/ * 2 /*bufs*/ * 2 /*heads*/ * 4096;   // 64 KB
constexpr int kXchgBytes = 2 /*bufs*/ * 4 /*warps*/ * 128 * 4;    // 4 KB
constexpr int kSmemBytes = kABytes + kWBytes + kOutBytes + kXchgBytes + 1024 + 256;
constexpr int kTmemCols = 512;                 // 2 buffers x (even 128 + odd 128)

__device__ __forceinline__ void named_bar_sync(int id, int n) {
    asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(n) : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint32_t bf16x2_max(uint32_t a, uint32_t b) {
    __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
    return *reinterpret_cast<uint32_t*>(&r);
}

__global__ void __launch_bounds__(kThreads, 1) stem_fused_kernel(const __grid_constant__ StemLaunch p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* a_sm = smem;                               // [buf][even|odd][128][128 B]
    uint8_t* w_sm = a_sm + kABytes;                     // [buf][128 co][128 B]
    uint8_t* out_sm = w_sm + kWBytes;                   // [warp][buf][head][32 px][128 B]
    float* xchg = reinterpret_cast<float*>(out_sm + kOutBytes);   // [buf][warp][128 ch]
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(xchg) + kXchgBytes);
    uint64_t* a_full = bars;            // [2] count 128 (builder threads)
    uint64_t* a_empty = bars + 2;       // [2] tcgen05.commit
    uint64_t* w_full = bars + 4;        // [2] TMA
    uint64_t* w_empty = bars + 6;       // [2] tcgen05.commit
    uint64_t* tmem_full = bars + 8;     // [2]
    uint64_t* tmem_empty = bars + 10;   // [2] count 4 (epilogue warps)
    uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 12);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.w_map);
        tma_prefetch_desc(&p.out_map);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&a_full[i], 128);
            mbar_init(&a_empty[i], 1);
            mbar_init(&w_full[i], 1);
            mbar_init(&w_empty[i], 1);
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 4);
        }
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc<kTmemCols>(tmem_base_slot);
    if (warp >= 1 && warp <= 4) {
        // chunk 7 of every A row is constant: {1, 1, 1, 0, 0, 0, 0, 0} (bias terms), written once
        const int i = (warp - 1) * 32 + lane;
        const uint4 ones = make_uint4(0x3F803F80u, 0x00003F80u, 0u, 0u);
        for (int t = 0; t < 4; ++t) *reinterpret_cast<uint4*>(a_sm + t * kTile + sw128_offset(i, 7)) = ones;
        fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_slot;

    const int units_per_group = p.B * kStrips;

    if (warp == 0) {
        // ------------------------------------------------------------------ weights TMA + UMMA issue
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(128, 128);
            uint32_t ui = 0, arow = 0;
            const int first = blockIdx.x;
            if (first < p.total_units) {
                mbar_expect_tx(&w_full[0], kTile);
                tma_load_2d(w_sm, &p.w_map, &w_full[0], 0, (first / units_per_group) * 128);
            }
            for (int u = first; u < p.total_units; u += gridDim.x, ++ui) {
                const int wb = ui & 1;
                const int un = u + gridDim.x;
                if (un < p.total_units) {                 // prefetch the next unit's weights into the other buffer
                    const int nb = wb ^ 1;
                    mbar_wait(&w_empty[nb], (((ui + 1) >> 1) & 1) ^ 1);
                    mbar_expect_tx(&w_full[nb], kTile);
                    tma_load_2d(w_sm + nb * kTile, &p.w_map, &w_full[nb], 0, (un / units_per_group) * 128);
                }
                mbar_wait(&w_full[wb], (ui >> 1) & 1);
                const uint32_t w_addr = smem_u32(w_sm + wb * kTile);
                for (int t = 0; t < kConvRowsPerUnit; ++t, ++arow) {
                    const int b = arow & 1;
                    const uint32_t ph = (arow >> 1) & 1;
                    mbar_wait(&a_full[b], ph);
                    mbar_wait(&tmem_empty[b], ph ^ 1);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(a_sm + b * 2 * kTile);
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        const uint64_t adesc = umma_desc_sw128(a_addr + half * kTile);
                        const uint64_t bdesc = umma_desc_sw128(w_addr);
                        const uint32_t d = tmem_base + b * 256 + half * 128;
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_bf16(d, adesc + 2 * k, bdesc + 2 * k, idesc, k ? 1u : 0u);
                    }
                    umma_commit(&a_empty[b]);
                    umma_commit(&tmem_full[b]);
                }
                umma_commit(&w_empty[wb]);
            }
        }
    } else if (warp <= 4) {
        // ------------------------------------------------------------------ builders: image -> A tiles
        const int i = (warp - 1) * 32 + lane;          // A row: even tile <-> conv x = 2i, odd tile <-> x = 2i+1
        uint32_t arow = 0;
        for (int u = blockIdx.x; u < p.total_units; u += gridDim.x) {
            const int r0 = u % units_per_group;
            const int img = r0 / kStrips;
            const int py0 = (r0 % kStrips) * kStripRows;
            const uint2* image = reinterpret_cast<const uint2*>(p.img + static_cast<size_t>(img) * 512 * 512);
            for (int t = 0; t < kConvRowsPerUnit; ++t, ++arow) {
                const int b = arow & 1;
                const int r = 2 * py0 - 1 + t;            // conv output row (may be -1: result is ignored)
                uint32_t w[7][6];
#pragma unroll
                for (int ky = 0; ky < 7; ++ky) {
                    const int iy = 2 * r + ky - 3;
                    const bool rowok = iy >= 0 && iy < 512;
                    const uint2* rp = image + iy * 128;   // 128 uint2 (4 pixels each) per image row
                    // pixels [4i-4, 4i+7] = uint2 index i-1, i, i+1
                    uint2 q0 = make_uint2(0u, 0u), q1 = q0, q2 = q0;
                    if (rowok) {
                        if (i > 0) q0 = __ldg(rp + i - 1);
                        q1 = __ldg(rp + i);
                        if (i < 127) q2 = __ldg(rp + i + 1);
                    }
                    w[ky][0] = q0.x; w[ky][1] = q0.y; w[ky][2] = q1.x; w[ky][3] = q1.y; w[ky][4] = q2.x; w[ky][5] = q2.y;
                }
                mbar_wait(&a_empty[b], ((arow >> 1) & 1) ^ 1);
                uint8_t* ae = a_sm + (b * 2 + 0) * kTile;
                uint8_t* ao = a_sm + (b * 2 + 1) * kTile;
#pragma unroll
                for (int ky = 0; ky < 7; ++ky) {
                    // even pixel x = 2i: image pixels [4i-3, 4i+4] = halves starting at the high half of word 0
                    const uint4 ce = make_uint4(__funnelshift_r(w[ky][0], w[ky][1], 16), __funnelshift_r(w[ky][1], w[ky][2], 16),
                                                __funnelshift_r(w[ky][2], w[ky][3], 16), __funnelshift_r(w[ky][3], w[ky][4], 16));
                    // odd pixel x = 2i+1: image pixels [4i-1, 4i+6]
                    const uint4 co = make_uint4(__funnelshift_r(w[ky][1], w[ky][2], 16), __funnelshift_r(w[ky][2], w[ky][3], 16),
                                                __funnelshift_r(w[ky][3], w[ky][4], 16), __funnelshift_r(w[ky][4], w[ky][5], 16));
                    *reinterpret_cast<uint4*>(ae + sw128_offset(i, ky)) = ce;
                    *reinterpret_cast<uint4*>(ao + sw128_offset(i, ky)) = co;
                }
                fence_proxy_async();
                mbar_arrive(&a_full[b]);
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue: pool + store
        const int q = warp & 3;                        // TMEM lane quarter; pooled px = q*32 + lane
        uint8_t* my_out = out_sm + q * (2 * 2 * 4096);
        uint32_t arow = 0, nemit = 0;
        uint32_t carry[64];                            // running vertical max, 128 channels as packed bf16
        for (int u = blockIdx.x; u < p.total_units; u += gridDim.x) {
            const int g = u / units_per_group;
            const int r0 = u % units_per_group;
            const int img = r0 / kStrips;
            const int py0 = (r0 % kStrips) * kStripRows;
            for (int t = 0; t < kConvRowsPerUnit; ++t, ++arow) {
                const int b = arow & 1;
                const int r = 2 * py0 - 1 + t;
                mbar_wait(&tmem_full[b], (arow >> 1) & 1);
                tc_fence_after();
                const uint32_t tbase = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + b * 256;
                float* xw = xchg + (b * 4 + q) * 128;
                // pre-pass: lane 31's odd-pixel values are the "x-1" neighbour of the next warp's lane 0
#pragma unroll 1
                for (int cb = 0; cb < 4; ++cb) {
                    uint32_t o[32];
                    tmem_ld32(tbase + 128 + cb * 32, o);
                    tmem_ld_wait();
                    if (lane == 31) {
#pragma unroll
                        for (int c = 0; c < 32; ++c) xw[cb * 32 + c] = __uint_as_float(o[c]);
                    }
                }
                named_bar_sync(1, 128);
                const float* xr = xchg + (b * 4 + (q > 0 ? q - 1 : 0)) * 128;
                const bool emit = (t >= 2) && ((t & 1) == 0);
                const int eb = nemit & 1;
                if (emit) {
                    if (lane == 0) tma_store_wait_read<1>();       // staging buffer `eb` has been read out
                    __syncwarp();
                }
#pragma unroll
                for (int cb = 0; cb < 4; ++cb) {
                    uint32_t e[32], o[32];
                    tmem_ld32(tbase + cb * 32, e);
                    tmem_ld32(tbase + 128 + cb * 32, o);
                    tmem_ld_wait();
                    if (cb == 3) {                                  // accumulators fully read: release the TMEM buffer
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tmem_empty[b]);
                    }
                    uint32_t hb[16];
#pragma unroll
                    for (int c = 0; c < 32; c += 2) {
                        float h2[2];
#pragma unroll
                        for (int s = 0; s < 2; ++s) {
                            const float ov = __uint_as_float(o[c + s]);
                            float op = __shfl_up_sync(0xffffffffu, ov, 1);
                            if (lane == 0) op = q > 0 ? xr[cb * 32 + c + s] : -INFINITY;
                            h2[s] = fmaxf(fmaxf(__uint_as_float(e[c + s]), ov), op);
                        }
                        hb[c >> 1] = pack_bf16(h2[0], h2[1]);
                    }
                    if (t == 0) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) carry[cb * 16 + j] = r < 0 ? 0xFF80FF80u /* -inf, -inf */ : hb[j];
                    } else if (!emit) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) carry[cb * 16 + j] = bf16x2_max(carry[cb * 16 + j], hb[j]);
                    } else {
                        // pooled row done: relu(max(carry, h)); this conv row also starts the next pooled row
                        uint8_t* stage = my_out + (eb * 2 + (cb >> 1)) * 4096;     // head = cb / 2
                        const uint32_t zero = 0u;
#pragma unroll
                        for (int j4 = 0; j4 < 4; ++j4) {
                            uint32_t v[4];
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj) {
                                const int j = j4 * 4 + jj;
                                v[jj] = bf16x2_max(bf16x2_max(carry[cb * 16 + j], hb[j]), zero);
                                carry[cb * 16 + j] = hb[j];
                            }
                            *reinterpret_cast<uint4*>(stage + sw128_offset(lane, (cb & 1) * 4 + j4)) =
                                make_uint4(v[0], v[1], v[2], v[3]);
                        }
                    }
                }
                if (emit) {
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        const int py = py0 + (t >> 1) - 1;
#pragma unroll
                        for (int hh = 0; hh < 2; ++hh) {
                            const int head = g * 2 + hh;
                            if (head < p.H) {
                                const int pix = ((head * p.B + img) * 128 + py) * 128 + q * 32;
                                tma_store_2d(&p.out_map, my_out + (eb * 2 + hh) * 4096, 0, pix);
                            }
                        }
                        tma_store_commit();
                    }
                    ++nemit;
                }
            }
        }
        if (lane == 0) tma_store_wait<0>();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<kTmemCols>(tmem_base);
}

}  // namespace

cudaError_t stem_fused_launch(const StemLaunch& p_in, int num_sms, cudaStream_t stream) {
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(stem_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    StemLaunch p = p_in;
    p.G = (p.H + 1) / 2;
    p.total_units = p.G * p.B * kStrips;
    const int grid = p.total_units < num_sms ? p.total_units : num_sms;
    stem_fused_kernel<<<grid, kThreads, kSmemBytes, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace sad
