// K3b': the row-stationary 3x3 / stride-1 convolution of the 64-channel 128x128 layer1 maps (conv_rows.cu) on CTA PAIRS:
// tcgen05 cta_group::2, UMMA 256x192x16, every instruction a full 192-column window.
//
// Why: layer1 is bound by the shared-memory operand feed of the tensor core (conv_rows.cu, block_rows.cu: 53% tensor
// pipe).  Per input row the 12 UMMAs of one CTA read the whole 72 KB weight set (B) next to 48 KB of A.  On a CTA pair
// each SM supplies its own 128-pixel A tile but only HALF of B (96 of the 192 rows), so an instruction costs 7 KB of
// shared-memory reads per SM instead of 10 KB, and the resident weights shrink to 36 KB per CTA.
//
// What makes the pair possible -- one B arrangement for EVERY instruction:
//   * the two CTAs of a pair take the two 64-row strips of ONE image (same head, same weights) and run in lock step;
//   * phantom rows: every input row is multiplied against all three vertical taps, also at the strip borders, where one
//     or two of the three destinations are output rows outside the strip -- they get an accumulator like real rows and
//     are thrown away (2 + 2 per 64-row strip: 6% more accumulator traffic, no extra instructions);
//   * aliased accumulator ring: TMEM holds 8 blocks of 64 columns; the window of input row g covers output rows
//     g, g+1, g+2 at blocks p, p+1, p+2 with p = g mod 6 <= 5, so a window never wraps.  An output row r with
//     r mod 6 >= 2 lives in block r mod 6; a row with r mod 6 = 0 (1) receives its first two (one) contributions in
//     block 6 (7) and the rest in block 0 (1), and the epilogue adds the two partial accumulators.  A strip consumes
//     66 = 6 * 11 input rows, so every unit starts at p = 0: which rows are split does not depend on where a unit sits
//     in a CTA's work list, and results stay independent of batch composition bit for bit.
//   (conv_rows.cu issues 2 of every 8 rows as a 128- plus a 64-column instruction because its 8-block ring wraps.)
//
// Pipeline per CTA: warp 0 TMA producer (own input rows and residual tiles; weights half), warp 1 UMMA issuer (leader CTA
// only), warps 2-5 epilogue (own 128 TMEM lanes).  Barriers follow conv_umma2.cu: loads of both CTAs complete on the
// LEADER's mbarrier, tcgen05.commit multicasts to both CTAs, epilogue warps of both CTAs release accumulator blocks on the
// leader.
//
// Replaces, for layer1.{0,1}.conv{1,2}, the conv2d+batch_norm(+add)+relu of timm's BasicBlock inside
// BinaryClassifier.forward (reference modular/source/inference_runner.py:49-51).
#include <cuda_runtime.h>

#include "conv_umma.h"
#include "ptx.cuh"

namespace sad {

namespace {

constexpr int kThreads = 192;
constexpr int kW = 128;
constexpr int kStripRows = 64;                // output rows per CTA and unit: the two strips of one image form a pair
constexpr int kInRows = kStripRows + 2;       // 66 = 6 * 11 input rows (windows) per unit
static_assert(kInRows % 6 == 0, "a unit must cover whole periods of the aliased ring");
constexpr int kRowBox = kW + 2;
constexpr int kRowBytes = kRowBox * 128;      // 16640 B written by one TMA box
constexpr int kSlotBytes = 17 * 1024;
constexpr int kRing = 6;
constexpr int kHalfRows = 96;                 // B rows per CTA and horizontal tap: half of [ky2 | ky1 | ky0] x 64 co
constexpr int kWHalfBytes = 3 * kHalfRows * 128;   // 36 KB
constexpr int kTileBytes = 128 * 128;
constexpr int kResRing = 3;
constexpr int kOutBytes = 4 * 2 * 4096;
constexpr int kSmemBytes = kWHalfBytes + kRing * kSlotBytes + kResRing * kTileBytes + kOutBytes + 1024 + 512;
constexpr int kTmemCols = 512;
static_assert(kSmemBytes <= 232448, "shared memory budget");

__device__ __forceinline__ void tmem_st32_zero(uint32_t taddr) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, "
        "%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};\n" ::"r"(taddr),
        "r"(0u)
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }

// first accumulator block of output row r (global row counter): rows with r mod 6 < 2 start in the alias blocks 6, 7
__device__ __forceinline__ uint32_t first_block(uint32_t r) {
    const uint32_t m = r % 6u;
    return m >= 2u ? m : 6u + m;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
    conv_rows2_kernel(const __grid_constant__ ConvLaunch p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* wsm = smem;                                    // [kx][96 rows of this CTA's half][128 B]
    uint8_t* ring = smem + kWHalfBytes;
    uint8_t* res_sm = ring + kRing * kSlotBytes;
    uint8_t* out_sm = res_sm + kResRing * kTileBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(out_sm + kOutBytes);
    uint64_t* in_full = bars;                   // [kRing]  leader: both CTAs' row loads
    uint64_t* in_empty = in_full + kRing;       // [kRing]  both (multicast commit)
    uint64_t* w_full = in_empty + kRing;        // [1]      leader
    uint64_t* w_empty = w_full + 1;             // [1]      both
    uint64_t* acc_full = w_empty + 1;           // [6]      both (multicast commit): output row complete
    uint64_t* acc_empty = acc_full + 6;         // [8]      leader: block read out and zeroed by 4 + 4 epilogue warps
    uint64_t* res_full = acc_empty + 8;         // [kResRing] local
    uint64_t* res_empty = res_full + kResRing;  // [kResRing] local
    uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(res_empty + kResRing);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1;
    const int n_pairs = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.a_map[0]);
        tma_prefetch_desc(&p.bh_map);
        tma_prefetch_desc(&p.out_map);
        tma_prefetch_desc(&p.res_map);
        for (int s = 0; s < kResRing; ++s) {
            mbar_init(&res_full[s], 1);
            mbar_init(&res_empty[s], 4);
        }
        for (int s = 0; s < kRing; ++s) {
            mbar_init(&in_full[s], 1);
            mbar_init(&in_empty[s], 1);
        }
        mbar_init(w_full, 1);
        mbar_init(w_empty, 1);
        for (int a = 0; a < 6; ++a) mbar_init(&acc_full[a], 1);
        for (int a = 0; a < 8; ++a) mbar_init(&acc_empty[a], 8);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_2sm<kTmemCols>(tmem_base_slot);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                             // barriers of both CTAs are initialised before any remote signal
    tc_fence_after();
    pdl_launch_dependents();
    pdl_wait();
    const uint32_t tmem_base = *tmem_base_slot;

    const int total_images = p.total_tiles;         // heads * imgs: one image (two strips) per pair and unit
    const int y0 = static_cast<int>(rank) * kStripRows;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer (both CTAs)
        if (lane == 0) {
            int cur_head = -1;
            uint32_t w_loads = 0, seq = 0, rseq = 0;
            const bool has_res = p.residual != nullptr;
            for (int img = pair; img < total_images; img += n_pairs) {
                const int head = img / p.imgs_per_head;
                if (head != cur_head) {
                    mbar_wait(w_empty, (w_loads & 1) ^ 1);          // the previous head's UMMAs have drained
                    if (rank == 0) mbar_expect_tx(w_full, 2 * kWHalfBytes);
                    // B rows in window order [ky = 2 | ky = 1 | ky = 0] x 64 co; this CTA holds rows rank*96 .. +95
                    for (int kx = 0; kx < 3; ++kx)
                        for (int b = 0; b < 3; ++b) {
                            const int r0 = static_cast<int>(rank) * kHalfRows + 32 * b;
                            const int ky = 2 - r0 / 64, co0 = r0 % 64;
                            tma_load_2d_2sm(wsm + (kx * kHalfRows + 32 * b) * 128, &p.bh_map, w_full, (ky * 3 + kx) * 64,
                                            head * 64 + co0);
                        }
                    ++w_loads;
                    cur_head = head;
                }
                for (int i = 0; i < kInRows; ++i, ++seq) {
                    const int slot = seq % kRing;
                    mbar_wait(&in_empty[slot], ((seq / kRing) & 1) ^ 1);
                    if (rank == 0) mbar_expect_tx(&in_full[slot], 2 * kRowBytes);
                    tma_load_4d_2sm(ring + slot * kSlotBytes, &p.a_map[0], &in_full[slot], 0, -1, y0 - 1 + i, img);
                    if (has_res && i >= 2) {                // residual tile of output row i-2 (complete after row i)
                        const int rs = rseq % kResRing;
                        mbar_wait(&res_empty[rs], ((rseq / kResRing) & 1) ^ 1);
                        mbar_expect_tx(&res_full[rs], kTileBytes);
                        tma_load_2d(res_sm + rs * kTileBytes, &p.res_map, &res_full[rs], 0, (img * kW + (y0 + i - 2)) * kW);
                        ++rseq;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ UMMA issuer (leader CTA only)
        if (lane == 0 && rank == 0) {
            const uint32_t w_addr = smem_u32(wsm);
            const uint32_t ring_addr = smem_u32(ring);
            constexpr uint32_t idesc = umma_idesc_bf16(256, 192);
            int cur_head = -1;
            uint32_t w_loads = 0;
            uint32_t g = 0;                         // input rows consumed = index of the window's first output row
            for (int img = pair; img < total_images; img += n_pairs) {
                const int head = img / p.imgs_per_head;
                if (head != cur_head) {
                    if (cur_head >= 0) umma_commit_2sm(w_empty);
                    mbar_wait(w_full, w_loads & 1);
                    ++w_loads;
                    cur_head = head;
                }
                for (int i = 0; i < kInRows; ++i, ++g) {
                    const uint32_t pblk = g % 6u, period = g / 6u;
                    // blocks this window starts using: the new output row g+2, and at p = 0 the rows g and g+1 that move
                    // from the alias blocks 6, 7 to blocks 0, 1.  Use n of a block waits for its release n (0 = zeroing).
                    mbar_wait(&acc_empty[first_block(g + 2)], ((g + 2) / 6u) & 1);
                    if (pblk == 0) {
                        mbar_wait(&acc_empty[0], period & 1);
                        mbar_wait(&acc_empty[1], period & 1);
                    }
                    const int slot = g % kRing;
                    mbar_wait(&in_full[slot], (g / kRing) & 1);
                    tc_fence_after();
                    const uint32_t row_addr = ring_addr + slot * kSlotBytes;
                    const uint32_t d_tmem = tmem_base + pblk * 64;
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        const uint64_t adesc = umma_desc_sw128(row_addr + kx * 128);
                        const uint64_t bdesc = umma_desc_sw128(w_addr + kx * kHalfRows * 128);
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_bf16_2sm(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, 1u);
                    }
                    umma_commit_2sm(&in_empty[slot]);
                    umma_commit_2sm(&acc_full[pblk]);       // output row g has all its contributions
                }
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue (warps 2..5 of both CTAs)
        const int quarter = warp & 3;
        const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
        for (int cidx = 0; cidx < kTmemCols; cidx += 32) tmem_st32_zero(tmem_base + lane_base + cidx);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0)
            for (int a = 0; a < 8; ++a) mbar_arrive_leader(&acc_empty[a]);

        uint8_t* my_out = out_sm + quarter * 2 * 4096;
        const bool has_res = p.residual != nullptr;
        const int prow = quarter * 32 + lane;
        uint32_t g = 0;                              // output-row counter, phantoms included
        uint32_t T = 0;                              // real output rows written (staging / residual ring position)
        for (int img = pair; img < total_images; img += n_pairs) {
            const int head = img / p.imgs_per_head;
            const float4* bias4 = reinterpret_cast<const float4*>(p.bias + head * 64);
            for (int i = 0; i < kInRows; ++i, ++g) {
                // row g of this unit is output row j = i - 2 of the strip; i = 0, 1 are the phantoms above the strip (the
                // two below it are rows i = 0, 1 of the NEXT unit: same counter, since 66 windows produce 66 completed rows)
                const uint32_t blk = g % 6u;
                mbar_wait(&acc_full[blk], (g / 6u) & 1);
                tc_fence_after();
                const uint32_t taddr = tmem_base + lane_base + blk * 64;
                const bool split = blk < 2;
                const uint32_t taddr2 = tmem_base + lane_base + (6 + blk) * 64;
                uint32_t v0[32], v1[32];
                const bool real = i >= 2;
                if (real) {
                    tmem_ld32(taddr, v0);
                    tmem_ld32(taddr + 32, v1);
                    tmem_ld_wait();
                    if (split) {                            // second partial accumulator (alias block)
                        uint32_t w0[32], w1[32];
                        tmem_ld32(taddr2, w0);
                        tmem_ld32(taddr2 + 32, w1);
                        tmem_ld_wait();
#pragma unroll
                        for (int q = 0; q < 32; ++q) {
                            v0[q] = __float_as_uint(__uint_as_float(w0[q]) + __uint_as_float(v0[q]));
                            v1[q] = __float_as_uint(__uint_as_float(w1[q]) + __uint_as_float(v1[q]));
                        }
                    }
                }
                tmem_st32_zero(taddr);
                tmem_st32_zero(taddr + 32);
                if (split) {
                    tmem_st32_zero(taddr2);
                    tmem_st32_zero(taddr2 + 32);
                }
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive_leader(&acc_empty[blk]);
                    if (split) mbar_arrive_leader(&acc_empty[6 + blk]);
                }
                if (!real) continue;
                const int j = i - 2;
                const int rs = T % kResRing;
                if (has_res) mbar_wait(&res_full[rs], (T / kResRing) & 1);
                if (lane == 0) tma_store_wait_read<1>();
                __syncwarp();
                uint8_t* stage = my_out + (T & 1) * 4096;
                const uint8_t* res_row = res_sm + rs * kTileBytes;
#pragma unroll
                for (int ch = 0; ch < 8; ++ch) {
                    const uint32_t* v = ch < 4 ? v0 : v1;
                    const int o = (ch & 3) * 8;
                    const float4 b0 = __ldg(bias4 + ch * 2), b1 = __ldg(bias4 + ch * 2 + 1);
                    float f[8] = {__uint_as_float(v[o + 0]) + b0.x, __uint_as_float(v[o + 1]) + b0.y,
                                  __uint_as_float(v[o + 2]) + b0.z, __uint_as_float(v[o + 3]) + b0.w,
                                  __uint_as_float(v[o + 4]) + b1.x, __uint_as_float(v[o + 5]) + b1.y,
                                  __uint_as_float(v[o + 6]) + b1.z, __uint_as_float(v[o + 7]) + b1.w};
                    if (has_res) {
                        const uint4 rr = ld_shared_v4(smem_u32(res_row) + sw128_offset(prow, ch));
                        const uint32_t rw[4] = {rr.x, rr.y, rr.z, rr.w};
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            f[2 * q] += act_lo(rw[q]);
                            f[2 * q + 1] += act_hi(rw[q]);
                        }
                    }
                    uint32_t pk[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        pk[q] = p.relu ? act_pack_relu(f[2 * q], f[2 * q + 1]) : act_pack(f[2 * q], f[2 * q + 1]);
                    st_shared_v4(smem_u32(stage) + sw128_offset(lane, ch), pk[0], pk[1], pk[2], pk[3]);
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    if (has_res) mbar_arrive(&res_empty[rs]);
                    tma_store_2d(&p.out_map, stage, 0, (img * kW + (y0 + j)) * kW + quarter * 32);
                    tma_store_commit();
                }
                ++T;
            }
        }
        if (lane == 0) tma_store_wait<0>();
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                             // the peer may still signal our barriers until here
    if (warp == 1) tmem_dealloc_2sm<kTmemCols>(tmem_base);
}

}  // namespace

// `p` as built for conv_rows_launch (a_map[0] with box {64, 130, 1, 1}); bh_map must be the weight view with box {64, 32}.
cudaError_t conv_rows2_launch(const ConvLaunch& p_in, int heads, int num_sms, cudaStream_t stream) {
    cudaError_t e = ensure_dynamic_smem<conv_rows2_kernel>(kSmemBytes);
    if (e != cudaSuccess) return e;
    ConvLaunch p = p_in;
    p.total_tiles = heads * p.imgs_per_head;        // images: the pair's two CTAs take the two strips of one image
    int pairs = num_sms / 2;
    if (p.total_tiles < pairs) pairs = p.total_tiles;
    if (pairs < 1) return cudaSuccess;
    return launch_pdl(conv_rows2_kernel, dim3(2 * pairs), dim3(kThreads), kSmemBytes, stream, p);
}

}  // namespace sad
