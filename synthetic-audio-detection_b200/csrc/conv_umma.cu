// K3: implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM).
//
// One launch = one convolution layer for ALL heads of the ensemble (grouped over heads):
//   out[img, oy, ox, co] = act( sum_{ky,kx,ci} in[img, oy*s+ky-p, ox*s+kx-p, ci] * w[head, co, ky, kx, ci]
//                               + bias[head, co] + residual[img, oy, ox, co] )
// with eval-mode BatchNorm already folded into w/bias (api.cu), activations NHWC bf16, fp32 accumulation.
// Replaces, per layer, the conv2d + batch_norm (+ add) + relu sequence that timm's ResNet runs inside
// BinaryClassifier.forward (reference modular/source/inference_runner.py:49-51).
//
// GEMM view:  M = 128 output pixels (a block of 128/Wo full output rows of one image),
//             N = N_TILE output channels, K = taps * Cin walked in blocks of 64 channels of one tap.
//   A tile (128 px x 64 ch, K-major, 128 B rows): ONE 4-D TMA box {64 ch, Wo, 128/Wo rows, 1 image} shifted by
//       the tap offset; out-of-bounds rows/columns (the conv zero padding) are zero-filled by TMA.
//       Stride-2 layers read through one of four "parity" tensor maps (a strided view of the input with
//       W/2 x H/2 pixels) so the box stays dense.
//   B tile (N_TILE x 64, K-major): 2-D TMA box from the packed weights [heads*Cout][taps*Cin].
//   Both land in SWIZZLE_128B layout and are consumed by tcgen05.mma (UMMA 128 x N_TILE x 16) straight from
//   shared memory; the fp32 accumulator lives in TMEM (double-buffered: 2 x N_TILE columns).
//
// Warp roles (192 threads, 1 CTA/SM, persistent over a static tile schedule):
//   warp 0 : TMA producer (one elected lane)      -- full/empty mbarrier ring of kStages stages
//   warp 1 : TMEM allocator + UMMA issuer (one lane), tcgen05.commit releases smem stages / signals epilogue
//   warps 2-5 : epilogue: tcgen05.ld -> +bias (+residual) -> ReLU -> bf16 -> global (NHWC)
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdlib>

#include "conv_umma.h"
#include "ptx.cuh"

namespace sad {

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                       // bf16 elements = 128 bytes = one swizzle row
constexpr int kABytes = kBlockM * kBlockK * 2;    // 16 KB

// MT = M tiles (of 128 pixels) that share one B (weight) stage.  With MT = 2 the same bytes in flight feed twice the
// tensor cycles; used for N_TILE = 128 (layer2), which was limited by stage bytes in flight over TMA latency.
template <int N_TILE, int MT>
struct Cfg {
    static constexpr int kBBytes = N_TILE * kBlockK * 2;
    static constexpr int kStageBytes = MT * kABytes + kBBytes;
    // epilogue staging: per epilogue warp kOutBufs buffers of [32 px][64 ch] bf16 (4 KB, SWIZZLE_128B) for TMA stores
    static constexpr int kAccCols = MT * N_TILE;                          // one accumulator set (MT tiles)
    static constexpr int kAccBufs = 2 * kAccCols <= 512 ? 2 : 1;          // double buffered when it fits TMEM
    static constexpr int kTmemCols = kAccBufs * kAccCols < 32 ? 32 : kAccBufs * kAccCols;   // 128 / 256 / 512
    static_assert(kTmemCols <= 512, "accumulators do not fit TMEM");
    // With a single accumulator set the MMA warp idles while it is drained, so the two M tiles are drained in parallel
    // by two epilogue warp groups (8 warps).
    static constexpr int kEpiGroups = (kAccBufs == 1 && MT == 2) ? 2 : 1;
    static constexpr int kThreadsCfg = 64 + 128 * kEpiGroups;
    static constexpr int kOutBufs = N_TILE == 256 ? 1 : 2;
    static constexpr int kOutBytes = 4 * kEpiGroups * kOutBufs * 4096;
    static constexpr int kBudget = 224 * 1024;
    static constexpr int kStages = ((kBudget - kOutBytes) / kStageBytes) > 8 ? 8 : ((kBudget - kOutBytes) / kStageBytes);
    static constexpr int kSmemBytes = kStages * kStageBytes + kOutBytes + 1024 /*align slack*/ + 256 /*barriers*/;
    static_assert(kSmemBytes <= 232448, "shared memory budget");
};

// TR ("transposed product", N_TILE = 128, MT = 2 only): the same stages -- two 128-pixel tiles and one 128-channel weight
// tile per K block -- are multiplied the other way round: A = weights (M = 128 output channels), B = the 256 pixels
// (N = 256), D[channel][pixel].  A tensor instruction then does 128x256x16 instead of 128x128x16; this part runs a
// 128-column instruction at ~61 % of the tensor peak whatever feeds it (on layer2 neither halving the operand reads with
// CTA pairs nor removing 8/9 of the A fetches moved it) and a 256-column one at ~76 %: +19 % on layer2.
// The price is an epilogue with channels on TMEM lanes: each thread owns one channel and 32 pixels per tcgen05.ld,
// writes 2-byte elements into a [32 px][32 ch] staging tile (a warp's 32 lanes fill 64 contiguous bytes per pixel) and
// the tile leaves by TMA.  A residual has no cheap place there (2-byte global loads: 848 instead of 1 023 TFLOP/s; TMA
// tiles one chunk ahead: 680), so api.cu feeds the identity branch through the tensor core instead: two extra K blocks
// of the block input against an identity weight tile (the mechanism of the folded downsample conv), and `residual` is
// null for this kernel.
template <int N_TILE, int MT, bool TR = false>
__global__ void __launch_bounds__(Cfg<N_TILE, MT>::kThreadsCfg, 1) conv_umma_kernel(const __grid_constant__ ConvLaunch p) {
    using C = Cfg<N_TILE, MT>;
    static_assert(!TR || (N_TILE == 128 && MT == 2), "transposed product: 128 channels x 256 pixels");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* tiles = smem;
    uint8_t* out_sm = smem + C::kStages * C::kStageBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(out_sm + C::kOutBytes);
    uint64_t* full_bar = bars;                      // [kStages]
    uint64_t* empty_bar = bars + C::kStages;        // [kStages]
    uint64_t* tmem_full = bars + 2 * C::kStages;    // [2]
    uint64_t* tmem_empty = tmem_full + 2;           // [2]
    uint64_t* res_bar = tmem_empty + 2;             // [4 warps][kOutBufs]  TR epilogue: residual tile landed (TMA)
    uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(res_bar + 4 * C::kOutBufs);
    static_assert((2 * C::kStages + 4 + 4 * C::kOutBufs) * 8 + 4 <= 256, "barrier block");

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < 4; ++i) tma_prefetch_desc(&p.a_map[i]);
        tma_prefetch_desc(&p.b_map);
        tma_prefetch_desc(&p.out_map);
        if (TR) tma_prefetch_desc(&p.out32_map);
        if (TR && p.residual) tma_prefetch_desc(&p.res32_map);
        if (p.k2_blocks) {
            tma_prefetch_desc(&p.a2_map);
            tma_prefetch_desc(&p.b2_map);
        }
        for (int s = 0; s < C::kStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full[a], 1);
            mbar_init(&tmem_empty[a], 4 * C::kEpiGroups);
        }
        if (TR)
            for (int i = 0; i < 4 * C::kOutBufs; ++i) mbar_init(&res_bar[i], 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<C::kTmemCols>(tmem_base_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_launch_dependents();                        // the next kernel in the stream may start its prologue on SMs we leave
    pdl_wait();                                     // our inputs (and buffers we overwrite) belong to the previous kernel until here
    const uint32_t tmem_base = *tmem_base_slot;

    const int taps = p.ksize * p.ksize;
    const int cblocks = p.Cin / kBlockK;
    const int ksteps = taps * cblocks + p.k2_blocks;    // + fused downsample branch (1x1, stride 2) K blocks
    const int m_groups = p.m_tiles_per_img / MT;              // groups of MT consecutive M tiles of one image
    const int tiles_per_head = p.imgs_per_head * m_groups * p.n_tiles;
    const int total_groups = p.total_tiles / MT;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total_groups; tile += gridDim.x) {
                const int head = tile / tiles_per_head;
                int r = tile - head * tiles_per_head;
                const int n_t = r % p.n_tiles;
                r /= p.n_tiles;
                const int m_t = (r % m_groups) * MT;
                const int img = r / m_groups;
                const int img_in = p.shared_input ? img : head * p.imgs_per_head + img;
                const int oy0 = m_t * p.rows_per_tile;
                const int wrow = head * p.Cout + n_t * N_TILE;
                for (int tap = 0; tap < taps; ++tap) {
                    const int offy = tap / p.ksize - p.pad;
                    const int offx = tap % p.ksize - p.pad;
                    int map = 0, x0 = offx, y0 = oy0 + offy;
                    if (p.stride == 2) {   // parity view: input pixel (2*h2+py, 2*w2+px)
                        map = ((offy & 1) << 1) | (offx & 1);
                        x0 = offx >> 1;    // arithmetic shift: floor(-1/2) = -1
                        y0 = oy0 + (offy >> 1);
                    }
                    for (int cb = 0; cb < cblocks; ++cb) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        uint8_t* a_dst = tiles + stage * C::kStageBytes;
                        uint8_t* b_dst = a_dst + MT * kABytes;
                        mbar_expect_tx(&full_bar[stage], C::kStageBytes);
#pragma unroll
                        for (int m = 0; m < MT; ++m)
                            tma_load_4d(a_dst + m * kABytes, &p.a_map[map], &full_bar[stage], cb * kBlockK, x0,
                                        y0 + m * p.rows_per_tile, img_in);
                        tma_load_2d(b_dst, &p.b_map, &full_bar[stage], tap * p.Cin + cb * kBlockK, wrow);
                        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
                    }
                }
                // fused downsample branch: out += x[2*oy, 2*ox, :] * w_ds  (parity-(0,0) view, no tap offset)
                for (int cb = 0; cb < p.k2_blocks; ++cb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* a_dst = tiles + stage * C::kStageBytes;
                    uint8_t* b_dst = a_dst + MT * kABytes;
                    mbar_expect_tx(&full_bar[stage], C::kStageBytes);
#pragma unroll
                    for (int m = 0; m < MT; ++m)
                        tma_load_4d(a_dst + m * kABytes, &p.a2_map, &full_bar[stage], cb * kBlockK, 0,
                                    oy0 + m * p.rows_per_tile, img_in);
                    tma_load_2d(b_dst, &p.b2_map, &full_bar[stage], cb * kBlockK, wrow);
                    if (++stage == C::kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ UMMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = TR ? umma_idesc_bf16(128, 256) : umma_idesc_bf16(kBlockM, N_TILE);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < total_groups; tile += gridDim.x, ++it) {
                const int acc = it % C::kAccBufs;
                mbar_wait(&tmem_empty[acc], ((it / C::kAccBufs) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * C::kAccCols;
                for (int ks = 0; ks < ksteps; ++ks) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(tiles + stage * C::kStageBytes);
                    const uint64_t bdesc = umma_desc_sw128(a_addr + MT * kABytes);
                    if constexpr (TR) {
                        // weights [128 co][64 k] as A, the two pixel tiles (contiguous: [256 px][64 k]) as B
                        const uint64_t pdesc = umma_desc_sw128(a_addr);
#pragma unroll
                        for (int k = 0; k < kBlockK / 16; ++k)
                            umma_bf16(d_tmem, bdesc + 2 * k, pdesc + 2 * k, idesc, (ks | k) != 0 ? 1u : 0u);
                    } else {
                    // k outer, m inner: consecutive instructions accumulate into DIFFERENT tiles (+0.7 % on layer2 over
                    // issuing each tile's four k-steps back to back)
#pragma unroll
                    for (int k = 0; k < kBlockK / 16; ++k) {
#pragma unroll
                        for (int m = 0; m < MT; ++m)
                            // advance 16 bf16 = 32 bytes along K inside the 128-byte swizzle row: +2 in >>4 units
                            umma_bf16(d_tmem + m * N_TILE, umma_desc_sw128(a_addr + m * kABytes) + 2 * k, bdesc + 2 * k, idesc,
                                      (ks | k) != 0 ? 1u : 0u);
                    }
                    }
                    umma_commit(&empty_bar[stage]);   // frees this smem stage once the MMAs have read it
                    if (++stage == C::kStages) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tmem_full[acc]);         // accumulator complete -> epilogue
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue (warps 2..5)
        // Per 64-channel block: TMEM -> registers -> +bias (+residual) -> ReLU -> bf16 -> swizzled smem staging ->
        // one TMA store of [32 px][64 ch] per warp (fully coalesced; a lane writing its own 128-byte pixel row
        // straight to global touches 32 lines per store instruction and made the epilogue the bottleneck).
        if constexpr (TR) {
            // ---- transposed product: TMEM lanes are output channels, columns are the 256 pixels of the two M tiles
            const int quarter = warp & 3;             // channels 32*quarter .. +31 of this 128-channel tile
            uint8_t* my_out = out_sm + quarter * C::kOutBufs * 4096;       // [32 px][32 ch] bf16 = 2 KB per buffer, no swizzle
            int it = 0;
            uint32_t nstore = 0;
            for (int tile = blockIdx.x; tile < total_groups; tile += gridDim.x, ++it) {
                const int head = tile / tiles_per_head;
                int r = tile - head * tiles_per_head;
                const int n_t = r % p.n_tiles;
                r /= p.n_tiles;
                const int m_t = (r % m_groups) * MT;
                const int img = r / m_groups;
                const int acc = it % C::kAccBufs;
                const int co = n_t * N_TILE + quarter * 32 + lane;                 // this thread's output channel
                const float bias = __ldg(p.bias + head * p.Cout + co);
                const long long pix_base = (static_cast<long long>(head) * p.imgs_per_head + img) * (p.m_tiles_per_img * kBlockM) +
                                           static_cast<long long>(m_t) * kBlockM;  // 256 consecutive pixels
                const bool has_res = p.residual != nullptr;
                constexpr int kChunks = MT * kBlockM / 32;
                // residual tile of chunk j lands in the upper 2 KB of staging buffer (nstore + j) % kOutBufs; the loads
                // run kOutBufs - 1 chunks ahead of their use (they do not depend on the accumulator)
                auto issue_res = [&](int j, uint32_t ns) {
                    const uint32_t b = ns % C::kOutBufs;
                    mbar_expect_tx(&res_bar[quarter * C::kOutBufs + b], 2048);
                    tma_load_2d(my_out + b * 4096 + 2048, &p.res32_map, &res_bar[quarter * C::kOutBufs + b],
                                n_t * N_TILE + quarter * 32, static_cast<int>(pix_base) + 32 * j);
                };
                if (has_res && lane == 0)
                    for (int j = 0; j < C::kOutBufs - 1 && j < kChunks; ++j) issue_res(j, nstore + j);
                mbar_wait(&tmem_full[acc], (it / C::kAccBufs) & 1);
                tc_fence_after();
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * C::kAccCols;
#pragma unroll 1
                for (int j = 0; j < kChunks; ++j, ++nstore) {
                    uint32_t v[32];
                    tmem_ld32(taddr + 32 * j, v);
                    tmem_ld_wait();
                    if (j == kChunks - 1) {            // accumulators are in registers: hand TMEM back
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tmem_empty[acc]);
                    }
                    if (lane == 0) {
                        tma_store_wait_read<C::kOutBufs - 1>();     // buffer nstore % kOutBufs (and the one ahead) may be rewritten
                        if (has_res && j + C::kOutBufs - 1 < kChunks) issue_res(j + C::kOutBufs - 1, nstore + C::kOutBufs - 1);
                    }
                    __syncwarp();
                    const uint32_t bsel = nstore % C::kOutBufs;
                    uint8_t* stage = my_out + bsel * 4096;
                    const uint32_t st_addr = smem_u32(stage) + lane * 2;          // [px][32 ch]: this thread's channel column
                    if (has_res) {
                        mbar_wait(&res_bar[quarter * C::kOutBufs + bsel], (nstore / C::kOutBufs) & 1);
#pragma unroll
                        for (int px = 0; px < 32; ++px) {
                            uint16_t rv;
                            asm volatile("ld.shared.u16 %0, [%1];" : "=h"(rv) : "r"(st_addr + 2048 + px * 64) : "memory");
                            v[px] = __float_as_uint(__uint_as_float(v[px]) + act_lo(static_cast<uint32_t>(rv)));
                        }
                    }
#pragma unroll
                    for (int px = 0; px < 32; px += 2) {
                        const float f0 = __uint_as_float(v[px]) + bias, f1 = __uint_as_float(v[px + 1]) + bias;
                        const uint32_t w = p.relu ? act_pack_relu(f0, f1) : act_pack(f0, f1);
                        st_shared_u16(st_addr + px * 64, w & 0xFFFFu);
                        st_shared_u16(st_addr + (px + 1) * 64, w >> 16);
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&p.out32_map, stage, n_t * N_TILE + quarter * 32, static_cast<int>(pix_base) + 32 * j);
                        tma_store_commit();
                    }
                }
            }
            if (lane == 0) tma_store_wait<0>();
        } else {
        const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
        const int row = quarter * 32 + lane;          // pixel inside the 128-pixel tile
        const int egroup = (warp - 2) >> 2;           // which epilogue warp group (0 unless kEpiGroups == 2)
        uint8_t* my_out = out_sm + (egroup * 4 + quarter) * C::kOutBufs * 4096;
        int it = 0;
        uint32_t nstore = 0;
        for (int tile = blockIdx.x; tile < total_groups; tile += gridDim.x, ++it) {
            const int head = tile / tiles_per_head;
            int r = tile - head * tiles_per_head;
            const int n_t = r % p.n_tiles;
            r /= p.n_tiles;
            const int m_t = (r % m_groups) * MT;
            const int img = r / m_groups;
            const int acc = it % C::kAccBufs;
            const int co0 = n_t * N_TILE;
            const float4* bias4 = reinterpret_cast<const float4*>(p.bias + head * p.Cout + co0);

            mbar_wait(&tmem_full[acc], (it / C::kAccBufs) & 1);
            tc_fence_after();
            const int m_lo = C::kEpiGroups == 2 ? egroup : 0;
            const int m_hi = C::kEpiGroups == 2 ? egroup + 1 : MT;
#pragma unroll 1
            for (int m = m_lo; m < m_hi; ++m) {
            const long long pix0 = (static_cast<long long>(head) * p.imgs_per_head + img) * (p.m_tiles_per_img * kBlockM) +
                                   (m_t + m) * kBlockM;
            const act_t* res = p.residual ? p.residual + (pix0 + row) * p.Cout + co0 : nullptr;
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * C::kAccCols + m * N_TILE;
#pragma unroll 1
            for (int c0 = 0; c0 < N_TILE; c0 += 64, ++nstore) {
                uint32_t v0[32], v1[32];
                tmem_ld32(taddr + c0, v0);
                tmem_ld32(taddr + c0 + 32, v1);
                uint4 rv[8];
                if (res) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) rv[q] = __ldg(reinterpret_cast<const uint4*>(res + c0) + q);
                }
                tmem_ld_wait();
                if (c0 + 64 >= N_TILE && m == m_hi - 1) {   // this group's accumulators are in registers: hand TMEM back
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tmem_empty[acc]);
                }
                if (lane == 0) tma_store_wait_read<C::kOutBufs - 1>();
                __syncwarp();
                uint8_t* stage = my_out + (nstore % C::kOutBufs) * 4096;
#pragma unroll
                for (int ch = 0; ch < 8; ++ch) {      // 8 channels = one 16-byte chunk
                    const uint32_t* v = ch < 4 ? v0 : v1;
                    const int o = (ch & 3) * 8;
                    const float4 b0 = __ldg(bias4 + (c0 >> 2) + ch * 2), b1 = __ldg(bias4 + (c0 >> 2) + ch * 2 + 1);
                    float f[8] = {__uint_as_float(v[o + 0]) + b0.x, __uint_as_float(v[o + 1]) + b0.y,
                                  __uint_as_float(v[o + 2]) + b0.z, __uint_as_float(v[o + 3]) + b0.w,
                                  __uint_as_float(v[o + 4]) + b1.x, __uint_as_float(v[o + 5]) + b1.y,
                                  __uint_as_float(v[o + 6]) + b1.z, __uint_as_float(v[o + 7]) + b1.w};
                    if (res) {
                        const uint32_t rw[4] = {rv[ch].x, rv[ch].y, rv[ch].z, rv[ch].w};
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
                    tma_store_2d(&p.out_map, stage, co0 + c0, static_cast<int>(pix0) + quarter * 32);
                    tma_store_commit();
                }
            }
            }
        }
        if (lane == 0) tma_store_wait<0>();
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<C::kTmemCols>(tmem_base);
}

template <int N_TILE, int MT, bool TR = false>
cudaError_t launch_t(const ConvLaunch& p, int num_sms, cudaStream_t stream) {
    using C = Cfg<N_TILE, MT>;
    cudaError_t e = ensure_dynamic_smem<conv_umma_kernel<N_TILE, MT, TR>>(C::kSmemBytes);
    if (e != cudaSuccess) return e;
    const int groups = p.total_tiles / MT;
    int grid = groups < num_sms ? groups : num_sms;
    return launch_pdl(conv_umma_kernel<N_TILE, MT, TR>, dim3(grid), dim3(C::kThreadsCfg), C::kSmemBytes, stream, p);
}

}  // namespace

int conv_n_tile(int Cout) { return Cout >= 256 ? 256 : Cout; }

cudaError_t conv_umma_launch(const ConvLaunch& p, int num_sms, cudaStream_t stream) {
    switch (p.n_tile) {
        case 64: return launch_t<64, 1>(p, num_sms, stream);
        case 128:
            if (p.m_tiles_per_img % 2 != 0) return launch_t<128, 1>(p, num_sms, stream);
            if (p.transposed && !p.residual) return launch_t<128, 2, true>(p, num_sms, stream);   // 256-column UMMA
            return launch_t<128, 2>(p, num_sms, stream);
        case 256: return (p.m_tiles_per_img % 2 == 0 && getenv("SAD_MT256") && atoi(getenv("SAD_MT256")) == 2)
                             ? launch_t<256, 2>(p, num_sms, stream) : launch_t<256, 1>(p, num_sms, stream);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace sad
