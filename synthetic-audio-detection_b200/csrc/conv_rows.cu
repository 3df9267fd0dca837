// K3b: row-stationary 3x3 / stride-1 convolution for the 64-channel, 128x128 layer1 maps (tcgen05 + TMEM).
//
// Why a second kernel: with Cout = 64 the generic kernel (conv_umma.cu) needs a fresh 16 KB A tile from L2 every
// 128 tensor cycles -- 9 TMA fetches of (almost) the same input rows per output tile -- and measured 27% of the
// tensor peak, exactly the unique-data L2->SM bandwidth (~35 B/clk/SM).  A first row-stationary version (one
// 128x64x16 UMMA per tap and k-step, 36 per output row) reached 47%: with N = 64 the tensor core is fed 6 KB of
// shared-memory operands per 32-cycle instruction and the operand fetch (~70 B/clk measured) becomes the limit.
// This version folds the three VERTICAL taps into N:
//
//   smem:  weights of the current head as 3 (kx) x [192 = 3 ky x 64 co][64 ci] bf16 (72 KB, SWIZZLE_128B, resident)
//          ring of halo'd input rows: one TMA box {64 ch, 130 px (x = -1..128), 1 row} = 130 x 128 B each; the
//          zero padding (x = -1, 128 and rows -1, 128) is TMA out-of-bounds fill.  Each row is fetched ONCE.
//   MMA :  input row i, horizontal tap kx:  A = 128 consecutive smem rows starting at pixel kx (descriptor start
//          shifted by kx*128 B), B = [ky*64+co][ci]  ->  D[x][ky*64+co] = contribution of input row i to output row
//          i-ky.  TMEM is a ring of 8 accumulators of 64 columns, output row T at slot (-T mod 8), so the three
//          destinations (T, T-1, T-2) are CONSECUTIVE columns and one UMMA 128x192x16 serves all three
//          (12 UMMAs per input row instead of 36; 120 KB instead of 216 KB of operand reads per output row).
//          All UMMAs accumulate; the epilogue hands a block back zeroed (tcgen05.st).
//   epilogue: TMEM -> +bias (+residual, TMA-loaded to smem) -> ReLU -> bf16 -> swizzled smem -> TMA store.
//   work:  unit = (head, image, strip of 32 output rows); persistent CTAs walk units round-robin; weights are
//          re-fetched only when the head changes.
//
// Replaces, for layer1.{0,1}.conv{1,2}, the conv2d+batch_norm(+add)+relu of timm's BasicBlock inside
// BinaryClassifier.forward (reference modular/source/inference_runner.py:49-51).
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "conv_umma.h"
#include "ptx.cuh"

namespace sad {

namespace {

constexpr int kThreads = 192;
constexpr int kW = 128;                       // output / input width and height of layer1
#ifndef SAD_ROWS_STRIP
#define SAD_ROWS_STRIP 32
#endif
constexpr int kStripRows = SAD_ROWS_STRIP;
constexpr int kStripsPerImg = kW / kStripRows;
constexpr int kInRows = kStripRows + 2;       // input rows per unit
constexpr int kRowBox = kW + 2;               // 130 pixels incl. the halo
constexpr int kRowBytes = kRowBox * 128;      // 16640 B written by one TMA box
constexpr int kSlotBytes = 17 * 1024;         // slot pitch (1024-aligned so the swizzle phase of pixel p is p mod 8)
constexpr int kRing = 4;
constexpr int kTapBytes = 64 * 128;           // one (kx,ky) block: [64 co][64 ci] bf16
constexpr int kWBytes = 9 * kTapBytes;        // 72 KB
constexpr int kTileBytes = 128 * 128;         // one output / residual tile: 128 px x 64 ch bf16
constexpr int kResRing = 3;
constexpr int kOutBytes = 4 * 2 * 4096;       // per epilogue warp: 2 staging buffers of 32 px x 128 B
constexpr int kSmemBytes = kWBytes + kRing * kSlotBytes + kResRing * kTileBytes + kOutBytes + 1024 + 512;
constexpr int kAccSlots = 8;
constexpr int kTmemCols = 512;                // 8 accumulators x 64 columns

// MEASURED on B200: the tensor core applies the 128B swizzle to the ABSOLUTE shared-memory address bits (like TMA
// does when it writes), so a descriptor whose start is shifted by whole 128-byte rows needs NO "matrix base offset"
// (bits 49-51 stay 0); setting it to (addr >> 7) & 7 gives wrong results.

__device__ __forceinline__ void tmem_st32_zero(uint32_t taddr) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, "
        "%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};\n" ::"r"(taddr),
        "r"(0u)
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }

__global__ void __launch_bounds__(kThreads, 1) conv_rows_kernel(const __grid_constant__ ConvLaunch p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* wsm = smem;                                    // [kx][ky][64 co][128 B]
    uint8_t* ring = smem + kWBytes;
    uint8_t* res_sm = ring + kRing * kSlotBytes;            // [kResRing][128 px][128 B] swizzled (TMA load)
    uint8_t* out_sm = res_sm + kResRing * kTileBytes;       // [4 warps][2][32 px][128 B] swizzled (TMA store)
    uint64_t* bars = reinterpret_cast<uint64_t*>(out_sm + kOutBytes);
    uint64_t* in_full = bars;                   // [kRing]
    uint64_t* in_empty = in_full + kRing;       // [kRing]
    uint64_t* w_full = in_empty + kRing;        // [1]
    uint64_t* w_empty = w_full + 1;             // [1]
    uint64_t* acc_full = w_empty + 1;           // [kAccSlots]  output row complete (tcgen05.commit)
    uint64_t* acc_empty = acc_full + kAccSlots; // [kAccSlots]  block read out and zeroed (4 epilogue warps)
    uint64_t* res_full = acc_empty + kAccSlots; // [kResRing]
    uint64_t* res_empty = res_full + kResRing;  // [kResRing]
    uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(res_empty + kResRing);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.a_map[0]);
        tma_prefetch_desc(&p.b_map);
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
        for (int a = 0; a < kAccSlots; ++a) {
            mbar_init(&acc_full[a], 1);
            mbar_init(&acc_empty[a], 4);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<kTmemCols>(tmem_base_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_launch_dependents();                        // the next kernel in the stream may start its prologue on SMs we leave
    pdl_wait();                                     // our inputs (and buffers we overwrite) belong to the previous kernel until here
    const uint32_t tmem_base = *tmem_base_slot;

    const int units_per_head = p.imgs_per_head * kStripsPerImg;
    const int total_units = p.total_tiles;      // heads * imgs * strips

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int cur_head = -1;
            uint32_t w_loads = 0;
            uint32_t seq = 0;                   // input rows loaded so far (ring position)
            uint32_t rseq = 0;                  // residual tiles loaded so far
            const bool has_res = p.residual != nullptr;
            for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
                const int head = u / units_per_head;
                const int r = u - head * units_per_head;
                const int img = head * p.imgs_per_head + r / kStripsPerImg;
                const int y0 = (r % kStripsPerImg) * kStripRows;
                if (head != cur_head) {
                    mbar_wait(w_empty, (w_loads & 1) ^ 1);      // previous head's MMAs have drained
                    mbar_expect_tx(w_full, kWBytes);
                    for (int kx = 0; kx < 3; ++kx)
                        for (int ky = 0; ky < 3; ++ky)
                            tma_load_2d(wsm + (kx * 3 + ky) * kTapBytes, &p.b_map, w_full, (ky * 3 + kx) * 64, head * 64);
                    ++w_loads;
                    cur_head = head;
                }
                for (int i = 0; i < kInRows; ++i, ++seq) {
                    const int slot = seq % kRing;
                    mbar_wait(&in_empty[slot], ((seq / kRing) & 1) ^ 1);
                    mbar_expect_tx(&in_full[slot], kRowBytes);
                    tma_load_4d(ring + slot * kSlotBytes, &p.a_map[0], &in_full[slot], 0, -1, y0 - 1 + i, img);
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
        // ------------------------------------------------------------------ UMMA issuer
        if (lane == 0) {
            const uint32_t w_addr = smem_u32(wsm);
            const uint32_t ring_addr = smem_u32(ring);
            int cur_head = -1;
            uint32_t w_loads = 0;
            uint32_t seq = 0;                   // input rows consumed so far
            uint32_t tbase = 0;                 // global index of this unit's output row 0
            for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
                const int head = u / units_per_head;
                if (head != cur_head) {
                    if (cur_head >= 0) umma_commit(w_empty);   // all MMAs that read the old weights are done
                    mbar_wait(w_full, w_loads & 1);
                    ++w_loads;
                    cur_head = head;
                }
                for (int i = 0; i < kInRows; ++i, ++seq) {
                    // input row i feeds output rows j = i - ky, ky in [ky_lo, ky_hi]
                    const int ky_lo = i >= kStripRows ? i - (kStripRows - 1) : 0;
                    const int ky_hi = i < 2 ? i : 2;
                    if (i < kStripRows) {       // output row j = i starts accumulating: its block must be free (zeroed)
                        const uint32_t T = tbase + i;
                        mbar_wait(&acc_empty[(0u - T) & 7u], (T >> 3) & 1);   // use n of a slot waits for empty-event n (0 = initial zeroing)
                    }
                    const int slot = seq % kRing;
                    mbar_wait(&in_full[slot], (seq / kRing) & 1);
                    tc_fence_after();
                    const uint32_t row_addr = ring_addr + slot * kSlotBytes;
                    // destination blocks: ky = ky_lo .. ky_hi  <->  output rows T_i - ky  <->  slots s0, s0+1, ... (mod 8)
                    const uint32_t s0 = (0u - (tbase + i - ky_lo)) & 7u;
                    const int nky = ky_hi - ky_lo + 1;
                    const int n_first = (s0 + nky <= 8u) ? nky : static_cast<int>(8u - s0);   // blocks before the ring wraps
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
                    if (i >= 2) umma_commit(&acc_full[(0u - (tbase + i - 2)) & 7u]);   // output row i-2 is complete
                }
                tbase += kStripRows;
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue (warps 2..5)
        // TMEM -> registers -> (+bias, +residual from smem, ReLU, bf16) -> swizzled smem -> TMA store.  Each warp owns
        // the 32 pixels of its TMEM lane quarter and two private 4 KB staging buffers, so no cross-warp barrier.
        const int quarter = warp & 3;
        const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
        // all accumulators start at zero (every UMMA accumulates)
        for (int cidx = 0; cidx < kTmemCols; cidx += 32) tmem_st32_zero(tmem_base + lane_base + cidx);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0)
            for (int a = 0; a < kAccSlots; ++a) mbar_arrive(&acc_empty[a]);

        uint8_t* my_out = out_sm + quarter * 2 * 4096;
        const bool has_res = p.residual != nullptr;
        uint32_t T = 0;                          // global output-row counter
        for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
            const int head = u / units_per_head;
            const int r = u - head * units_per_head;
            const int img = head * p.imgs_per_head + r / kStripsPerImg;
            const int y0 = (r % kStripsPerImg) * kStripRows;
            const float4* bias4 = reinterpret_cast<const float4*>(p.bias + head * 64);
            for (int j = 0; j < kStripRows; ++j, ++T) {
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
                if (lane == 0) mbar_arrive(&acc_empty[slot]);             // block is zero again: next user may start
                const int rs = T % kResRing;
                if (has_res) mbar_wait(&res_full[rs], (T / kResRing) & 1);
                if (lane == 0) tma_store_wait_read<1>();                  // staging buffer (T & 1) is free again
                __syncwarp();
                uint8_t* stage = my_out + (T & 1) * 4096;
                const uint8_t* res_row = res_sm + rs * kTileBytes;
                const int prow = quarter * 32 + lane;                     // pixel row inside the 128-px tile
#pragma unroll
                for (int ch = 0; ch < 8; ++ch) {                          // 8 channels = one 16-byte chunk
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
                    for (int q = 0; q < 4; ++q) {
                        pk[q] = p.relu ? act_pack_relu(f[2 * q], f[2 * q + 1]) : act_pack(f[2 * q], f[2 * q + 1]);
                    }
                    st_shared_v4(smem_u32(stage) + sw128_offset(lane, ch), pk[0], pk[1], pk[2], pk[3]);
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    if (has_res) mbar_arrive(&res_empty[rs]);
                    tma_store_2d(&p.out_map, stage, 0, (img * kW + (y0 + j)) * kW + quarter * 32);
                    tma_store_commit();
                }
            }
        }
        if (lane == 0) tma_store_wait<0>();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<kTmemCols>(tmem_base);
}

}  // namespace

// `p` as built for the generic kernel, except: a_map[0] must have box {64, 130, 1, 1}; total_tiles is recomputed
// here as the number of (head, image, strip) units.
cudaError_t conv_rows_launch(const ConvLaunch& p_in, int heads, int num_sms, cudaStream_t stream) {
    cudaError_t e = ensure_dynamic_smem<conv_rows_kernel>(kSmemBytes);
    if (e != cudaSuccess) return e;
    ConvLaunch p = p_in;
    p.total_tiles = heads * p.imgs_per_head * kStripsPerImg;
    const int grid = p.total_tiles < num_sms ? p.total_tiles : num_sms;
    return launch_pdl(conv_rows_kernel, dim3(grid), dim3(kThreads), kSmemBytes, stream, p);
}

}  // namespace sad
