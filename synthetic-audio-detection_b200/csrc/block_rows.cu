// K3c: one whole layer1 BasicBlock per launch on CTA PAIRS (cluster of 2), tcgen05 + TMEM + distributed shared memory.
//
//   out = relu( conv2( relu(conv1(x) + b1) ) + b2 + x )        64 -> 64 -> 64 channels, 3x3 / stride 1, 128x128 maps
//
// Why: run as two launches of conv_rows.cu the block moves 8 GB per chunk through HBM (x in, mid out, mid in, x in
// again as the residual, out) and conv2 sits at 86% of the measured copy bandwidth.  Here the intermediate never
// leaves the chip:
//
//   even CTA ("rank 0")  conv1: the row-stationary pipeline of conv_rows.cu (resident weights, ring of halo'd input
//                        rows, vertical taps folded into N = 192, TMEM ring of 8 accumulators).  Its epilogue adds the
//                        bias, applies ReLU, rounds to bf16, stages the row in its own shared memory already in the
//                        SWIZZLE_128B image of the PEER's input-row ring slot (phase = absolute address bits, what
//                        TMA would have produced) and hands it to the DMA engine: cp.async.bulk shared::cta ->
//                        shared::cluster, 4 KB per epilogue warp, bytes counted on the peer's "row full" mbarrier.
//                        (Measured: per-lane st.shared::cluster + fence.proxy.async + arrive.release spent 70% of
//                        the epilogue in the fence, 3.1 ms per block; st.async with complete_tx per 16 B, 2.56 ms:
//                        ~3 clk per remote store; two launches take 1.64 ms.)
//   odd CTA ("rank 1")   conv2: the same pipeline, its input rows arriving from the peer instead of from TMA; the
//                        residual tile (a row of x the peer fetched microseconds ago: an L2 hit) comes by TMA; bias +
//                        residual + ReLU, TMA store.  A relay thread turns "UMMAs that read ring slot s have
//                        retired" (tcgen05.commit, local) into a credit on the PEER's barrier so rank 0 may overwrite it.
//
// HBM traffic per block: x once in, out once out (3.2 GB per chunk instead of 8).  The two CTAs run the same number
// of tensor instructions per row, so the pair is balanced; a unit is (head, image, strip of kS output rows) and rank 0
// computes kS+2 intermediate rows for it (the two halo rows are recomputed by the neighbouring unit: 2/kS extra).
// Intermediate rows outside the image are conv2's zero padding and are written as zeros.
//
// Numerics are those of the two-launch path bit for bit: same bf16 rounding of the intermediate, same accumulation
// order (tests/test_gpu_conv.py compares the two).
//
// Replaces layer1.{0,1} of timm's ResNet inside BinaryClassifier.forward
// (reference modular/source/inference_runner.py:49-51).
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "conv_umma.h"
#include "ptx.cuh"

namespace sad {

namespace {

constexpr int kThreads = 224;                 // warp 0 TMA, warp 1 UMMA, warps 2-5 epilogue, warp 6 credit relay (rank 1)
constexpr int kW = 128;
#ifndef SAD_BLOCK_STRIP
#define SAD_BLOCK_STRIP 128   // whole images: no halo rows recomputed by conv1 (64-row strips: 2 of 66, measured -0.8 % whole path)
#endif
constexpr int kS = SAD_BLOCK_STRIP;           // output rows per unit
constexpr int kStripsPerImg = kW / kS;
constexpr int kRowBox = kW + 2;               // 130 pixels incl. the halo
constexpr int kRowBytes = kRowBox * 128;      // 16640 B written by one TMA box
constexpr int kSlotBytes = 17 * 1024;         // slot pitch (1024-aligned: swizzle phase of pixel row r is r mod 8)
#ifndef SAD_BLOCK_RING0
#define SAD_BLOCK_RING0 5
#endif
constexpr int kRing0 = SAD_BLOCK_RING0;       // rank 0: ring of input rows (TMA)
#ifndef SAD_BLOCK_MID
#define SAD_BLOCK_MID 4
#endif
constexpr int kMid = SAD_BLOCK_MID;           // rank 1: ring of intermediate rows (written by the peer), same offset
constexpr int kTapBytes = 64 * 128;
constexpr int kWBytes = 9 * kTapBytes;        // 72 KB
constexpr int kTileBytes = 128 * 128;
constexpr int kResRing = kMid > 4 ? 2 : 3;
constexpr int kOutBytes = 4 * 2 * 4096;
constexpr int kRank0Bytes = kRing0 * kSlotBytes + kMid * kTileBytes;   // + one staging tile per peer ring slot
constexpr int kRank1Bytes = kMid * kSlotBytes + kResRing * kTileBytes + kOutBytes;
constexpr int kBarOffset = kWBytes + (kRank0Bytes > kRank1Bytes ? kRank0Bytes : kRank1Bytes);
constexpr int kSmemBytes = kBarOffset + 1024 + 512;
constexpr int kMaxRing = kRing0 > kMid ? kRing0 : kMid;
constexpr int kAccSlots = 8;
constexpr int kTmemCols = 512;
static_assert(kSmemBytes <= 232448, "shared memory budget");
static_assert(kW % kS == 0, "strip must divide the image height");

__device__ __forceinline__ void tmem_st32_zero(uint32_t taddr) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, "
        "%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};\n" ::"r"(taddr),
        "r"(0u)
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
    block_rows_kernel(const __grid_constant__ ConvLaunch p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* wsm = smem;                                    // [kx][ky][64 co][128 B]
    uint8_t* ring = smem + kWBytes;                         // rank 0: kRing0 slots; rank 1: kMid slots (same offset)
    uint8_t* stage_sm = ring + kRing0 * kSlotBytes;         // rank 0 only: [kMid][128 px][128 B], image of peer slot rows 1..128
    uint8_t* res_sm = ring + kMid * kSlotBytes;             // rank 1 only
    uint8_t* out_sm = res_sm + kResRing * kTileBytes;       // rank 1 only
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kBarOffset);
    uint64_t* in_full = bars;                      // [kMaxRing]  row present (rank 0: TMA bytes; rank 1: the peer's st.async bytes)
    uint64_t* in_empty = in_full + kMaxRing;       // [kMaxRing]  UMMAs that read the row have retired (tcgen05.commit)
    uint64_t* w_full = in_empty + kMaxRing;        // [1]
    uint64_t* w_empty = w_full + 1;                // [1]
    uint64_t* acc_full = w_empty + 1;              // [kAccSlots]
    uint64_t* acc_empty = acc_full + kAccSlots;    // [kAccSlots]
    uint64_t* res_full = acc_empty + kAccSlots;    // [kResRing]
    uint64_t* res_empty = res_full + kResRing;     // [kResRing]
    uint64_t* mid_credit = res_empty + kResRing;   // [kMid]  rank 0's copy is used: the peer's ring slot may be overwritten
    uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(mid_credit + kMid);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int n_out = rank == 0 ? kS + 2 : kS;    // rows this CTA produces per unit
    const int n_in = n_out + 2;                   // rows it consumes
    const int ring_n = rank == 0 ? kRing0 : kMid;
    const CUtensorMap* wmap = rank == 0 ? &p.b_map : &p.b2_map;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.a_map[0]);
        tma_prefetch_desc(wmap);
        tma_prefetch_desc(&p.out_map);
        tma_prefetch_desc(&p.res_map);
        for (int s = 0; s < kResRing; ++s) {
            mbar_init(&res_full[s], 1);
            mbar_init(&res_empty[s], 4);
        }
        for (int s = 0; s < kMaxRing; ++s) {
            mbar_init(&in_full[s], 1);              // one expect_tx arrive; the bytes come from TMA (rank 0) or st.async (rank 1)
            mbar_init(&in_empty[s], 1);
        }
        for (int s = 0; s < kMid; ++s) mbar_init(&mid_credit[s], 1);
        mbar_init(w_full, 1);
        mbar_init(w_empty, 1);
        for (int a = 0; a < kAccSlots; ++a) {
            mbar_init(&acc_full[a], 1);
            mbar_init(&acc_empty[a], 4);
        }
        fence_barrier_init();
    }
    if (rank == 1) {
        // the halo pixels (x = -1, 128) of every intermediate-row slot stay zero for the whole kernel: the peer only
        // ever writes pixels 0..127 (slot rows 1..128)
        for (int i = threadIdx.x; i < kMid * kSlotBytes / 16; i += kThreads)
            reinterpret_cast<uint4*>(ring)[i] = make_uint4(0, 0, 0, 0);
        fence_proxy_async();
    }
    if (warp == 1) tmem_alloc<kTmemCols>(tmem_base_slot);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                           // both CTAs' barriers are initialised before any remote arrive / store
    tc_fence_after();
    pdl_launch_dependents();                        // the next kernel in the stream may start its prologue on SMs we leave
    pdl_wait();                                     // our inputs (and buffers we overwrite) belong to the previous kernel until here
    const uint32_t tmem_base = *tmem_base_slot;

    const int units_per_head = p.imgs_per_head * kStripsPerImg;
    const int total_units = p.total_tiles;        // heads * imgs * strips
    const int cid = blockIdx.x >> 1;
    const int n_clusters = gridDim.x >> 1;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int cur_head = -1;
            uint32_t w_loads = 0;
            uint32_t seq = 0;                     // rank 0: input rows loaded; rank 1: residual tiles loaded
            for (int u = cid; u < total_units; u += n_clusters) {
                const int head = u / units_per_head;
                const int r = u - head * units_per_head;
                const int img = head * p.imgs_per_head + r / kStripsPerImg;
                const int y0 = (r % kStripsPerImg) * kS;
                if (head != cur_head) {
                    mbar_wait(w_empty, (w_loads & 1) ^ 1);
                    mbar_expect_tx(w_full, kWBytes);
                    for (int kx = 0; kx < 3; ++kx)
                        for (int ky = 0; ky < 3; ++ky)
                            tma_load_2d(wsm + (kx * 3 + ky) * kTapBytes, wmap, w_full, (ky * 3 + kx) * 64, head * 64);
                    ++w_loads;
                    cur_head = head;
                }
                if (rank == 0) {
                    for (int i = 0; i < kS + 4; ++i, ++seq) {              // x rows y0-2 .. y0+kS+1 (zero outside the image)
                        const int slot = seq % kRing0;
                        mbar_wait(&in_empty[slot], ((seq / kRing0) & 1) ^ 1);
                        mbar_expect_tx(&in_full[slot], kRowBytes);
                        tma_load_4d(ring + slot * kSlotBytes, &p.a_map[0], &in_full[slot], 0, -1, y0 - 2 + i, img);
                    }
                } else {
                    for (int j = 0; j < kS; ++j, ++seq) {                  // residual = x row y0+j
                        const int rs = seq % kResRing;
                        mbar_wait(&res_empty[rs], ((seq / kResRing) & 1) ^ 1);
                        mbar_expect_tx(&res_full[rs], kTileBytes);
                        tma_load_2d(res_sm + rs * kTileBytes, &p.res_map, &res_full[rs], 0, (img * kW + (y0 + j)) * kW);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ UMMA issuer (both ranks)
        if (lane == 0) {
            const uint32_t w_addr = smem_u32(wsm);
            const uint32_t ring_addr = smem_u32(ring);
            int cur_head = -1;
            uint32_t w_loads = 0;
            uint32_t seq = 0;                     // rows consumed so far
            uint32_t tbase = 0;                   // global index of this unit's first produced row
            for (int u = cid; u < total_units; u += n_clusters) {
                const int head = u / units_per_head;
                if (head != cur_head) {
                    if (cur_head >= 0) umma_commit(w_empty);
                    mbar_wait(w_full, w_loads & 1);
                    ++w_loads;
                    cur_head = head;
                }
                for (int i = 0; i < n_in; ++i, ++seq) {
                    // consumed row i feeds produced rows j = i - ky, ky in [ky_lo, ky_hi]
                    const int ky_lo = i >= n_out ? i - (n_out - 1) : 0;
                    const int ky_hi = i < 2 ? i : 2;
                    if (i < n_out) {
                        const uint32_t T = tbase + i;
                        mbar_wait(&acc_empty[(0u - T) & 7u], (T >> 3) & 1);
                    }
                    const int slot = seq % ring_n;
                    // rank 1: the row is written by the peer's bulk copies (async proxy, complete_tx on this barrier) --
                    // the same visibility contract as a TMA load, so a CTA-scope wait and no proxy fence.  (An
                    // acquire.cluster wait compiles to TRYWAIT + CCTL.IVALL: it invalidated L1 once per row and cost 13%.)
                    if (rank == 1) mbar_expect_tx(&in_full[slot], kTileBytes);
                    mbar_wait(&in_full[slot], (seq / ring_n) & 1);
                    tc_fence_after();
                    const uint32_t row_addr = ring_addr + slot * kSlotBytes;
                    const uint32_t s0 = (0u - (tbase + i - ky_lo)) & 7u;
                    const int nky = ky_hi - ky_lo + 1;
                    const int n_first = (s0 + nky <= 8u) ? nky : static_cast<int>(8u - s0);
#pragma unroll 1
                    for (int seg = 0; seg < 2; ++seg) {
                        const int nb = seg == 0 ? n_first : nky - n_first;
                        if (nb == 0) break;
                        const int kyb = seg == 0 ? ky_lo : ky_lo + n_first;
                        const uint32_t d_tmem = tmem_base + (seg == 0 ? s0 : 0u) * 64;
                        const uint32_t idesc = umma_idesc_bf16(128, 64 * nb);
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx) {
                            const uint64_t adesc = umma_desc_sw128(row_addr + kx * 128);
                            const uint64_t bdesc = umma_desc_sw128(w_addr + (kx * 3 + kyb) * kTapBytes);
#pragma unroll
                            for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, 1u);
                        }
                    }
                    umma_commit(&in_empty[slot]);
                    if (i >= 2) umma_commit(&acc_full[(0u - (tbase + i - 2)) & 7u]);
                }
                tbase += n_out;
            }
        }
    } else if (warp == 6) {
        // ------------------------------------------------------------------ credit relay (rank 1): slot consumed -> peer
        if (rank == 1 && lane == 0) {
            const uint32_t credit0 = mapa_u32(smem_u32(mid_credit), 0);
            uint32_t seq = 0;
            for (int u = cid; u < total_units; u += n_clusters)
                for (int i = 0; i < kS + 2; ++i, ++seq) {
                    const int slot = seq % kMid;
                    mbar_wait(&in_empty[slot], (seq / kMid) & 1);
                    mbar_arrive_cluster(credit0 + slot * 8);
                }
        }
    } else {
        // ------------------------------------------------------------------ epilogue (warps 2..5)
        const int quarter = warp & 3;
        const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
        for (int cidx = 0; cidx < kTmemCols; cidx += 32) tmem_st32_zero(tmem_base + lane_base + cidx);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0)
            for (int a = 0; a < kAccSlots; ++a) mbar_arrive(&acc_empty[a]);

        const int prow = quarter * 32 + lane;                             // pixel of this lane inside the row
        uint32_t T = 0;                                                   // global produced-row counter
        if (rank == 0) {
            // conv1: +bias, ReLU, bf16 -> the peer's ring slot (pixel x lives in slot row x+1)
            const uint32_t peer_ring = mapa_u32(smem_u32(ring), 1);
            const uint32_t peer_full = mapa_u32(smem_u32(in_full), 1);
            for (int u = cid; u < total_units; u += n_clusters) {
                const int head = u / units_per_head;
                const int r = u - head * units_per_head;
                const int y0 = (r % kStripsPerImg) * kS;
                const float4* bias4 = reinterpret_cast<const float4*>(p.bias + head * 64);
                for (int j = 0; j < kS + 2; ++j, ++T) {
                    const uint32_t slot = (0u - T) & 7u;
                    mbar_wait(&acc_full[slot], (T >> 3) & 1);
                    tc_fence_after();
                    const uint32_t taddr = tmem_base + lane_base + slot * 64;
                    uint32_t v0[32], v1[32];
                    tmem_ld32(taddr, v0);
                    tmem_ld32(taddr + 32, v1);
                    tmem_ld_wait();
                    tmem_st32_zero(taddr);
                    tmem_st32_zero(taddr + 32);
                    tmem_st_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_empty[slot]);
                    const int my = y0 - 1 + j;                            // image row of this intermediate row
                    const bool inside = my >= 0 && my < kW;               // outside: conv2's zero padding
                    const uint32_t ms = T % kMid;
                    mbar_wait(&mid_credit[ms], ((T / kMid) & 1) ^ 1);     // the peer's UMMAs have retired the slot's previous row
                    uint8_t* stg = stage_sm + ms * kTileBytes;            // free: the credit implies its last copy has landed
#pragma unroll
                    for (int ch = 0; ch < 8; ++ch) {
                        const uint32_t* v = ch < 4 ? v0 : v1;
                        const int o = (ch & 3) * 8;
                        const float4 b0 = __ldg(bias4 + ch * 2), b1 = __ldg(bias4 + ch * 2 + 1);
                        const float f[8] = {__uint_as_float(v[o + 0]) + b0.x, __uint_as_float(v[o + 1]) + b0.y,
                                            __uint_as_float(v[o + 2]) + b0.z, __uint_as_float(v[o + 3]) + b0.w,
                                            __uint_as_float(v[o + 4]) + b1.x, __uint_as_float(v[o + 5]) + b1.y,
                                            __uint_as_float(v[o + 6]) + b1.z, __uint_as_float(v[o + 7]) + b1.w};
                        uint32_t pk[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            pk[q] = inside ? act_pack_relu(f[2 * q], f[2 * q + 1]) : 0u;
                        }
                        // pixel p goes to peer slot row p+1: same bytes within the row, swizzle phase (p+1) & 7
                        st_shared_v4(smem_u32(stg) + prow * 128 + ((ch ^ ((prow + 1) & 7)) << 4), pk[0], pk[1], pk[2], pk[3]);
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0)                                        // this warp's 32 pixels: 4 KB by the DMA engine
                        bulk_copy_to_peer(peer_ring + ms * kSlotBytes + 128 + quarter * 4096,
                                          smem_u32(stg) + quarter * 4096, 4096, peer_full + ms * 8);
                }
            }
        } else {
            // conv2: +bias +residual, ReLU, bf16 -> swizzled staging -> TMA store
            uint8_t* my_out = out_sm + quarter * 2 * 4096;
            for (int u = cid; u < total_units; u += n_clusters) {
                const int head = u / units_per_head;
                const int r = u - head * units_per_head;
                const int img = head * p.imgs_per_head + r / kStripsPerImg;
                const int y0 = (r % kStripsPerImg) * kS;
                const float4* bias4 = reinterpret_cast<const float4*>(p.bias2 + head * 64);
                for (int j = 0; j < kS; ++j, ++T) {
                    const uint32_t slot = (0u - T) & 7u;
                    mbar_wait(&acc_full[slot], (T >> 3) & 1);
                    tc_fence_after();
                    const uint32_t taddr = tmem_base + lane_base + slot * 64;
                    uint32_t v0[32], v1[32];
                    tmem_ld32(taddr, v0);
                    tmem_ld32(taddr + 32, v1);
                    tmem_ld_wait();
                    tmem_st32_zero(taddr);
                    tmem_st32_zero(taddr + 32);
                    tmem_st_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_empty[slot]);
                    const int rs = T % kResRing;
                    mbar_wait(&res_full[rs], (T / kResRing) & 1);
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
                        const uint4 rr = ld_shared_v4(smem_u32(res_row) + sw128_offset(prow, ch));
                        const uint32_t rw[4] = {rr.x, rr.y, rr.z, rr.w};
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            f[2 * q] += act_lo(rw[q]);
                            f[2 * q + 1] += act_hi(rw[q]);
                        }
                        uint32_t pk[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            pk[q] = act_pack_relu(f[2 * q], f[2 * q + 1]);
                        }
                        st_shared_v4(smem_u32(stage) + sw128_offset(lane, ch), pk[0], pk[1], pk[2], pk[3]);
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive(&res_empty[rs]);
                        tma_store_2d(&p.out_map, stage, 0, (img * kW + (y0 + j)) * kW + quarter * 32);
                        tma_store_commit();
                    }
                }
            }
            if (lane == 0) tma_store_wait<0>();
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                           // nobody exits while its peer may still store to / arrive on its smem
    if (warp == 1) tmem_dealloc<kTmemCols>(tmem_base);
}

}  // namespace

// `p`: a_map[0] = x with box {64, 130, 1, 1}; b_map / bias = conv1, b2_map / bias2 = conv2; res_map = x as 128-px tiles;
// out_map = block output.  One cluster (2 CTAs) per pair of SMs.
cudaError_t block_rows_launch(const ConvLaunch& p_in, int heads, int num_sms, cudaStream_t stream) {
    cudaError_t e = ensure_dynamic_smem<block_rows_kernel>(kSmemBytes);
    if (e != cudaSuccess) return e;
    ConvLaunch p = p_in;
    p.total_tiles = heads * p.imgs_per_head * kStripsPerImg;
    int clusters = num_sms / 2;
    if (p.total_tiles < clusters) clusters = p.total_tiles;
    if (clusters < 1) return cudaSuccess;
    return launch_pdl(block_rows_kernel, dim3(2 * clusters), dim3(kThreads), kSmemBytes, stream, p);
}

}  // namespace sad
