// K2: fused stem for the single-channel log-mel image:
//     conv 7x7 / stride 2 / pad 3 (1 -> 64 channels per head, BN folded) + ReLU + maxpool 3x3 / stride 2 / pad 1
// for TWO heads at a time, never materialising the 64 x 256 x 256 conv output (8 MB per head and segment).
// Replaces conv1 / bn1 / act1 / maxpool of timm's ResNet inside BinaryClassifier.forward (reference
// modular/source/inference_runner.py:49-51); the three identical input channels (:173) are folded into one by
// summing conv1's weights over Cin (api.cu).
//
// GEMM view per conv-output row r:   D[co][x] = sum_k W[co][k] * P[x][k]
//   M = 128 output channels (2 heads x 64)  -> TMEM LANES are channels
//   N = 256 conv pixels of the row          -> TMEM COLUMNS are pixels
//   K = 64: chunk ky (16 B) of pixel x = the 8 image pixels [2x-3, 2x+4] of image row 2r+ky-3 (7 taps + one
//       zero-weight slot); chunk 7 = {1,1,1,0,...} against {bias_hi, bias_mid, bias_lo} so the folded BN shift is
//       added by the tensor core in fp32 (bias split into three bf16 terms, exact to 24 bits).
// With channels on lanes, both pooling directions are per-thread register work: horizontal max over columns
// 2p-1, 2p, 2p+1, vertical max carried across conv rows as packed bf16 -- no shuffles, no cross-warp exchange.
// (A first version had pixels on lanes and was epilogue-bound: 12 TMEM round trips, 128 shuffles and a block
// barrier per conv row; ncu showed the builders and the MMA warp spinning on the epilogue's barriers.)
//   builders (4 warps): read the bf16 image (L2 resident) with 8-byte loads, funnel-shift, write the [256 px][128 B]
//                       pixel tile into SWIZZLE_128B shared memory; double buffered.
//   MMA (1 thread)    : 4 x tcgen05.mma 128x256x16 per conv row into TMEM (2 buffers x 256 columns).
//   epilogue (8 warps): per thread = one channel x one half of the row (128 conv px): 4 x tcgen05.ld of 32 columns (16 warps measured slower)
//                       (software-pipelined against the pooling math), pooling, ReLU, bf16; every second conv row the
//                       pooled row is transposed through swizzled smem and written with TMA stores
//                       ([128 px][64 ch] per head).
// Work unit = (image, head pair, strip of 32 pooled rows); persistent CTAs, round-robin.
#include <cuda_runtime.h>

#include <type_traits>

#include "conv_umma.h"
#include "ptx.cuh"
#include "stem_fused.h"

namespace sad {

namespace {

constexpr int kEpiGroups = 2;                  // epilogue warp groups; each owns 256/kEpiGroups conv pixels of a row
constexpr int kEpiThreads = 128 * kEpiGroups;
constexpr int kGroupCols = 256 / kEpiGroups;   // conv pixels (TMEM columns) per group
constexpr int kChunks = kGroupCols / 32;       // tcgen05.ld chunks of 32 columns per group and row
constexpr int kThreads = 32 + 128 + kEpiThreads;   // warp 0: MMA + TMA; warps 1-4: builders; then the epilogue warps
#ifndef SAD_STEM_STRIP
#define SAD_STEM_STRIP 32
#endif
constexpr int kStripRows = SAD_STEM_STRIP;     // pooled rows per unit
constexpr int kStrips = 128 / kStripRows;
constexpr int kConvRowsPerUnit = 2 * kStripRows + 1;
constexpr int kPixTile = 256 * 128;            // pixel tile: 256 rows x 128 B
constexpr int kWTile = 128 * 128;              // weight tile: 128 co x 64 k
constexpr int kPBytes = 2 * kPixTile;          // 64 KB (double buffered)
constexpr int kWBytes = 2 * kWTile;            // 32 KB (double buffered)
constexpr int kHeadTile = 128 * 128;           // staging: [128 px][64 ch] bf16 per head
constexpr int kOutBytes = 2 /*bufs*/ * 2 /*heads*/ * kHeadTile;   // 64 KB
constexpr int kSmemBytes = kPBytes + kWBytes + kOutBytes + 1024 + 256;
constexpr int kTmemCols = 512;                 // 2 buffers x 256 pixel columns

__device__ __forceinline__ void named_bar_sync(int id, int n) {
    asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(n) : "memory");
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {   // FMNMX3: one instruction on sm_100
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

// Pixel x of a conv row sits in row pix_row(x) of the pixel tile (= TMEM column pix_row(x) of the accumulator): inside
// every block of 16 pixels the 8 even ones come first, then the 8 odd ones.  A builder thread writes the two pixels
// 2i and 2i+1; with pixel x in row x the eight lanes of a quarter warp hit rows 0, 2, 4, .. 14 -- only four distinct
// values of (row & 7), i.e. four distinct 16-byte columns of the 128-byte swizzle: every tile store was a 2-way bank
// conflict (ncu, round 2: 26 M of the kernel's 34 M conflict wavefronts).  De-interleaved, the quarter warp's even
// pixels land in rows 16b .. 16b+7 and its odd ones in 16b+8 .. 16b+15: eight distinct columns, no conflict.  The
// epilogue reads accumulator columns by compile-time index, so the permutation costs it nothing.
__host__ __device__ constexpr int pix_row(int x) { return (x & ~15) | ((x & 1) << 3) | ((x & 15) >> 1); }

__global__ void __launch_bounds__(kThreads, 1) stem_fused_kernel(const __grid_constant__ StemLaunch p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* p_sm = smem;                               // [buf][256 px][128 B]
    uint8_t* w_sm = p_sm + kPBytes;                     // [buf][128 co][128 B]
    uint8_t* out_sm = w_sm + kWBytes;                   // [buf][head][128 px][128 B]
    uint64_t* bars = reinterpret_cast<uint64_t*>(out_sm + kOutBytes);
    uint64_t* p_full = bars;            // [2] count 128 (builder threads)
    uint64_t* p_empty = bars + 2;       // [2] tcgen05.commit
    uint64_t* w_full = bars + 4;        // [2] TMA
    uint64_t* w_empty = bars + 6;       // [2] tcgen05.commit
    uint64_t* tmem_full = bars + 8;     // [2]
    uint64_t* tmem_empty = bars + 10;   // [2] count 4 * kEpiGroups (epilogue warps)
    uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 12);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.w_map);
        tma_prefetch_desc(&p.out_map);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&p_full[i], 128);
            mbar_init(&p_empty[i], 1);
            mbar_init(&w_full[i], 1);
            mbar_init(&w_empty[i], 1);
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 4 * kEpiGroups);
        }
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc<kTmemCols>(tmem_base_slot);
    if (warp >= 1 && warp <= 4) {
        // chunk 7 of every pixel row is constant: {1, 1, 1, 0, 0, 0, 0, 0} (bias terms), written once
        const int i = (warp - 1) * 32 + lane;
        const uint4 ones = make_uint4(kActOne | (kActOne << 16), kActOne, 0u, 0u);
        for (int b = 0; b < 2; ++b) {
            *reinterpret_cast<uint4*>(p_sm + b * kPixTile + sw128_offset(2 * i, 7)) = ones;
            *reinterpret_cast<uint4*>(p_sm + b * kPixTile + sw128_offset(2 * i + 1, 7)) = ones;
        }
        fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_launch_dependents();                        // the next kernel in the stream may start its prologue on SMs we leave
    pdl_wait();                                     // our inputs (and buffers we overwrite) belong to the previous kernel until here
    const uint32_t tmem_base = *tmem_base_slot;

    const int units_per_group = p.B * kStrips;

    if (warp == 0) {
        // ------------------------------------------------------------------ weights TMA + UMMA issue
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(128, 256);
            uint32_t ui = 0, arow = 0;
            const int first = blockIdx.x;
            if (first < p.total_units) {
                mbar_expect_tx(&w_full[0], kWTile);
                tma_load_2d(w_sm, &p.w_map, &w_full[0], 0, (first / units_per_group) * 128);
            }
            for (int u = first; u < p.total_units; u += gridDim.x, ++ui) {
                const int wb = ui & 1;
                const int un = u + gridDim.x;
                if (un < p.total_units) {                 // prefetch the next unit's weights into the other buffer
                    const int nb = wb ^ 1;
                    mbar_wait(&w_empty[nb], (((ui + 1) >> 1) & 1) ^ 1);
                    mbar_expect_tx(&w_full[nb], kWTile);
                    tma_load_2d(w_sm + nb * kWTile, &p.w_map, &w_full[nb], 0, (un / units_per_group) * 128);
                }
                mbar_wait(&w_full[wb], (ui >> 1) & 1);
                const uint64_t adesc = umma_desc_sw128(smem_u32(w_sm + wb * kWTile));
                for (int t = 0; t < kConvRowsPerUnit; ++t, ++arow) {
                    const int b = arow & 1;
                    const uint32_t ph = (arow >> 1) & 1;
                    mbar_wait(&p_full[b], ph);
                    mbar_wait(&tmem_empty[b], ph ^ 1);
                    tc_fence_after();
                    const uint64_t bdesc = umma_desc_sw128(smem_u32(p_sm + b * kPixTile));
                    const uint32_t d = tmem_base + b * 256;
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16(d, adesc + 2 * k, bdesc + 2 * k, idesc, k ? 1u : 0u);
                    umma_commit(&p_empty[b]);
                    umma_commit(&tmem_full[b]);
                }
                umma_commit(&w_empty[wb]);
            }
        }
    } else if (warp <= 4) {
        // ------------------------------------------------------------------ builders: image -> pixel tile
        const int i = (warp - 1) * 32 + lane;          // builds conv pixels x = 2i and 2i+1
        uint32_t arow = 0;
        uint32_t off_e[7], off_o[7];                    // swizzled chunk offsets of this thread's two pixel rows
#pragma unroll
        for (int ky = 0; ky < 7; ++ky) {
            off_e[ky] = smem_u32(p_sm) + sw128_offset(pix_row(2 * i), ky);
            off_o[ky] = smem_u32(p_sm) + sw128_offset(pix_row(2 * i + 1), ky);
        }
        for (int u = blockIdx.x; u < p.total_units; u += gridDim.x) {
            const int r0 = u % units_per_group;
            const int img = r0 / kStrips;
            const int py0 = (r0 % kStrips) * kStripRows;
            const uint2* image = reinterpret_cast<const uint2*>(p.img + static_cast<size_t>(img) * 512 * 512);
            // Consecutive conv rows share 5 of their 7 image rows (stride 2): chunk ky of row r+1 is chunk ky+2 of row r.
            // The shifted 16-byte chunks stay in registers and roll down by two per row; only the two new image rows are
            // loaded -- one conv row AHEAD, so their L2 latency hides behind this row's buffer wait and stores
            // (ncu, round 2: the UMMA thread spent its time waiting for the builders, and the builders 45% of theirs
            // on the first use of freshly loaded pixels).
            // ce / co are ROTATING files of seven registers: chunk ky of conv row t lives in slot (ky + 2t) % 7, so the two
            // new image rows of row t overwrite the slots rows 0 and 1 of row t-1 left and nothing is moved -- the loop over
            // rows is unrolled by seven with the slot numbers as compile-time constants (ncu, round 2: the version that
            // shifted the arrays down by two per row executed 208 register moves of its 322 instructions per row).
            uint4 ce[7], co[7];                           // even pixel x = 2i / odd pixel x = 2i+1
            uint32_t nw[2][6];                            // raw words of the next row's two new image rows
            auto load_row = [&](int iy, uint32_t (&w)[6]) {
                uint2 q0 = make_uint2(0u, 0u), q1 = q0, q2 = q0;
                if (iy >= 0 && iy < 512) {
                    const uint2* rp = image + iy * 128;   // 128 uint2 (4 pixels each) per image row
                    if (i > 0) q0 = __ldg(rp + i - 1);    // pixels [4i-4, 4i+7] = uint2 index i-1, i, i+1
                    q1 = __ldg(rp + i);
                    if (i < 127) q2 = __ldg(rp + i + 1);
                }
                w[0] = q0.x; w[1] = q0.y; w[2] = q1.x; w[3] = q1.y; w[4] = q2.x; w[5] = q2.y;
            };
            auto shift_row = [&](const uint32_t (&w)[6], uint4& e, uint4& o) {
                // even pixel x = 2i: image pixels [4i-3, 4i+4] = halves starting at the high half of word 0
                e = make_uint4(__funnelshift_r(w[0], w[1], 16), __funnelshift_r(w[1], w[2], 16),
                               __funnelshift_r(w[2], w[3], 16), __funnelshift_r(w[3], w[4], 16));
                // odd pixel x = 2i+1: image pixels [4i-1, 4i+6]
                o = make_uint4(__funnelshift_r(w[1], w[2], 16), __funnelshift_r(w[2], w[3], 16),
                               __funnelshift_r(w[3], w[4], 16), __funnelshift_r(w[4], w[5], 16));
            };
            {   // first conv row of the unit: all seven image rows (slot ky = chunk ky)
                const int r = 2 * py0 - 1;
#pragma unroll
                for (int ky = 0; ky < 7; ++ky) {
                    uint32_t w[6];
                    load_row(2 * r + ky - 3, w);
                    shift_row(w, ce[ky], co[ky]);
                }
                load_row(2 * (r + 1) + 2, nw[0]);
                load_row(2 * (r + 1) + 3, nw[1]);
            }
            // one conv row; U = t % 7 fixes the slot of every chunk at compile time
            auto conv_row = [&](auto uc, int t) {
                constexpr int U = decltype(uc)::value;
                const int b = arow & 1;
                const int r = 2 * py0 - 1 + t;            // conv output row (may be -1: result is ignored)
                if (t > 0) {
                    shift_row(nw[0], ce[(5 + 2 * U) % 7], co[(5 + 2 * U) % 7]);
                    shift_row(nw[1], ce[(6 + 2 * U) % 7], co[(6 + 2 * U) % 7]);
                    if (t + 1 < kConvRowsPerUnit) {       // prefetch for row r+1: image rows 2(r+1)+2, 2(r+1)+3
                        load_row(2 * (r + 1) + 2, nw[0]);
                        load_row(2 * (r + 1) + 3, nw[1]);
                    }
                }
                mbar_wait(&p_empty[b], ((arow >> 1) & 1) ^ 1);
                const uint32_t tile = b * kPixTile;
#pragma unroll
                for (int ky = 0; ky < 7; ++ky) {
                    const uint4& e = ce[(ky + 2 * U) % 7];
                    const uint4& o = co[(ky + 2 * U) % 7];
                    st_shared_v4(off_e[ky] + tile, e.x, e.y, e.z, e.w);
                    st_shared_v4(off_o[ky] + tile, o.x, o.y, o.z, o.w);
                }
                fence_proxy_async();
                mbar_arrive(&p_full[b]);
                ++arow;
            };
            for (int t = 0; t < kConvRowsPerUnit; t += 7) {
                conv_row(std::integral_constant<int, 0>{}, t);
                if (t + 1 < kConvRowsPerUnit) conv_row(std::integral_constant<int, 1>{}, t + 1);
                if (t + 2 < kConvRowsPerUnit) conv_row(std::integral_constant<int, 2>{}, t + 2);
                if (t + 3 < kConvRowsPerUnit) conv_row(std::integral_constant<int, 3>{}, t + 3);
                if (t + 4 < kConvRowsPerUnit) conv_row(std::integral_constant<int, 4>{}, t + 4);
                if (t + 5 < kConvRowsPerUnit) conv_row(std::integral_constant<int, 5>{}, t + 5);
                if (t + 6 < kConvRowsPerUnit) conv_row(std::integral_constant<int, 6>{}, t + 6);
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue: pool + transpose + store
        const int q = warp & 3;                        // TMEM lane quarter
        const int half = (warp - 5) >> 2;              // column group: conv px [half*kGroupCols, +kGroupCols)
        const int c = q * 32 + lane;                   // channel within the head pair (lane of the accumulator)
        const int hh = c >> 6;                         // head within the pair
        const int cc = c & 63;                         // channel within the head
        const int et = threadIdx.x - 5 * 32;           // 0 .. kEpiThreads-1 within the epilogue warps
        uint32_t arow = 0, nemit = 0;
        uint32_t carry[kGroupCols / 4];                // running vertical max: kGroupCols/2 pooled px as packed bf16 pairs
        for (int u = blockIdx.x; u < p.total_units; u += gridDim.x) {
            const int g = u / units_per_group;
            const int r0 = u % units_per_group;
            const int img = r0 / kStrips;
            const int py0 = (r0 % kStrips) * kStripRows;
            for (int t = 0; t < kConvRowsPerUnit; ++t, ++arow) {
                const int b = arow & 1;
                const int r = 2 * py0 - 1 + t;
                const bool emit = (t >= 2) && ((t & 1) == 0);
                const int eb = nemit & 1;
                if (emit) {
                    if (et == 0) tma_store_wait_read<1>();         // staging buffer `eb` has been read out
                    named_bar_sync(1, kEpiThreads);
                }
                mbar_wait(&tmem_full[b], (arow >> 1) & 1);
                tc_fence_after();
                const uint32_t tbase = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + b * 256 + half * kGroupCols;
                // staging tile of this thread's head: shared-space addresses, one base per value of (px & 7) so that
                // every store below is [base + immediate] (a generic-pointer version compiled to ST + 4 address ops each)
                const uint32_t stage = smem_u32(out_sm) + (eb * 2 + hh) * kHeadTile + (cc & 7) * 2 + half * (kGroupCols / 2) * 128;
                uint32_t sbase[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) sbase[k] = stage + ((((cc >> 3) ^ k) & 7) << 4);
                uint32_t va[32], vb[32];
                float prev = -INFINITY;                            // conv pixel 2p-1 of the first pooled px
                if (half) {
                    uint32_t pv;
                    tmem_ld1(tbase - 1, pv);                       // the conv px just left of this group's columns
                    tmem_ld32(tbase, va);
                    tmem_ld_wait_dep(va);
                    asm volatile("" : "+r"(pv));
                    prev = __uint_as_float(pv);
                } else {
                    tmem_ld32(tbase, va);
                    tmem_ld_wait_dep(va);
                }
#pragma unroll
                for (int cb = 0; cb < kChunks; ++cb) {             // 32 conv pixels -> 16 pooled pixels
                    const uint32_t* v = (cb & 1) ? vb : va;
                    if (cb < kChunks - 1) {                          // next chunk in flight while this one is pooled
                        if (cb & 1) tmem_ld32(tbase + (cb + 1) * 32, va);
                        else tmem_ld32(tbase + (cb + 1) * 32, vb);
                    }
                    uint32_t hb[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float a0 = __uint_as_float(v[pix_row(4 * j)]), a1 = __uint_as_float(v[pix_row(4 * j + 1)]);
                        const float a2 = __uint_as_float(v[pix_row(4 * j + 2)]), a3 = __uint_as_float(v[pix_row(4 * j + 3)]);
                        const float h0 = fmax3(prev, a0, a1);             // pooled px 2j   : conv 2p-1, 2p, 2p+1
                        const float h1 = fmax3(a1, a2, a3);               // pooled px 2j+1
                        prev = a3;
                        hb[j] = act_pack_relu(h0, h1);
                    }
                    if (t == 0) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) carry[cb * 8 + j] = r < 0 ? 0u /* max identity after ReLU */ : hb[j];
                    } else if (!emit) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) carry[cb * 8 + j] = act_max2(carry[cb * 8 + j], hb[j]);
                    } else {
                        // pooled row done: max(carry, h) (ReLU already applied); this conv row also starts the next pooled row.
                        // Transpose through smem: this thread owns channel cc, pooled px half*64 + 16cb .. +15.
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const uint32_t o = act_max2(carry[cb * 8 + j], hb[j]);
                            carry[cb * 8 + j] = hb[j];
                            const int px = cb * 16 + 2 * j;                 // within this group's 64 pooled px (a multiple of 8 apart)
                            st_shared_u16(sbase[px & 7] + px * 128, o & 0xFFFFu);
                            st_shared_u16(sbase[(px + 1) & 7] + (px + 1) * 128, o >> 16);
                        }
                    }
                    if (cb < kChunks - 1) {
                        if (cb & 1) tmem_ld_wait_dep(va);
                        else tmem_ld_wait_dep(vb);
                    }
                    if (cb == kChunks - 2) {                        // last chunk is in registers: release the TMEM buffer
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tmem_empty[b]);
                    }
                }
                if (emit) {
                    fence_proxy_async();
                    named_bar_sync(1, kEpiThreads);
                    if (et == 0) {
                        const int py = py0 + (t >> 1) - 1;
#pragma unroll
                        for (int h2 = 0; h2 < 2; ++h2) {
                            const int head = g * 2 + h2;
                            if (head < p.H) {
                                const int pix = ((head * p.B + img) * 128 + py) * 128;
                                tma_store_2d(&p.out_map, out_sm + (eb * 2 + h2) * kHeadTile, 0, pix);
                            }
                        }
                        tma_store_commit();
                    }
                    ++nemit;
                }
            }
        }
        if (et == 0) tma_store_wait<0>();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<kTmemCols>(tmem_base);
}

}  // namespace

cudaError_t stem_fused_launch(const StemLaunch& p_in, int num_sms, cudaStream_t stream) {
    cudaError_t e = ensure_dynamic_smem<stem_fused_kernel>(kSmemBytes);
    if (e != cudaSuccess) return e;
    StemLaunch p = p_in;
    p.G = (p.H + 1) / 2;
    p.total_units = p.G * p.B * kStrips;
    const int grid = p.total_units < num_sms ? p.total_units : num_sms;
    return launch_pdl(stem_fused_kernel, dim3(grid), dim3(kThreads), kSmemBytes, stream, p);
}

}  // namespace sad
