// K3 (2-CTA variant): the implicit-GEMM convolution of conv_umma.cu on CTA PAIRS (tcgen05 cta_group::2).
//
// A cluster of two CTAs (one TPC) computes a 256-pixel x N_TILE tile: CTA r of the pair owns M tile 2g+r (its own
// 128-pixel A boxes, its own 128 accumulator rows in its own TMEM) and loads only rows [r*N/2, (r+1)*N/2) of each
// weight stage.  One thread of the even CTA issues UMMA 256 x N_TILE x 16 for both.  Per SM a K step then moves
// 16 KB + N_TILE*64 B instead of 16 KB + N_TILE*128 B through TMA, and the tensor core reads 4 KB + N_TILE*16 B
// of operands per instruction from each SM's shared memory instead of 4 KB + N_TILE*32 B -- the two limits measured on
// the 1-CTA kernel for N = 128 (layer2: 61-63% tensor pipe) and N = 256 (layers 3-4: 80-89%).
//
// Barriers: full[s]  (leader's copy only; count 1 = the leader's arrive.expect_tx of BOTH CTAs' bytes; both CTAs'
//                     TMA loads complete_tx on it),
//           empty[s], tmem_full[a] (one copy per CTA, signalled by multicast tcgen05.commit),
//           tmem_empty[a] (leader's copy; count 8 = the four epilogue warps of each CTA).
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdio>

#include "conv_umma.h"
#include "ptx.cuh"

namespace sad {

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kABytes = kBlockM * kBlockK * 2;    // 16 KB
constexpr int kThreads2 = 192;

template <int N_TILE, int MT>
struct Cfg2 {
    static constexpr int kBHalfBytes = (N_TILE / 2) * kBlockK * 2;
    static constexpr int kStageBytes = MT * kABytes + kBHalfBytes;        // per CTA: MT M tiles share one weight stage
    static constexpr int kOutBufs = 2;
    static constexpr int kOutBytes = 4 * kOutBufs * 4096;
    static constexpr int kBudget = 224 * 1024;
    static constexpr int kStages = ((kBudget - kOutBytes) / kStageBytes) > 8 ? 8 : ((kBudget - kOutBytes) / kStageBytes);
    static constexpr int kAccCols = MT * N_TILE;
    static constexpr int kTmemCols = 2 * kAccCols;                        // double-buffered accumulators (own rows)
    static_assert(kTmemCols <= 512, "accumulators do not fit TMEM");
    static constexpr int kSmemBytes = kStages * kStageBytes + kOutBytes + 1024 + 256;
    static_assert(kSmemBytes <= 232448, "shared memory budget");
};

template <int N_TILE, int MT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads2, 1)
    conv_umma2_kernel(const __grid_constant__ ConvLaunch p) {
    using C = Cfg2<N_TILE, MT>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* tiles = smem;
    uint8_t* out_sm = smem + C::kStages * C::kStageBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(out_sm + C::kOutBytes);
    uint64_t* full_bar = bars;                      // [kStages]   (used in the leader CTA)
    uint64_t* empty_bar = bars + C::kStages;        // [kStages]
    uint64_t* tmem_full = bars + 2 * C::kStages;    // [2]
    uint64_t* tmem_empty = tmem_full + 2;           // [2]         (used in the leader CTA)
    uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();        // 0 = leader (even SM of the pair)
    const int pair = blockIdx.x >> 1;
    const int n_pairs = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < 4; ++i) tma_prefetch_desc(&p.a_map[i]);
        tma_prefetch_desc(&p.bh_map);
        tma_prefetch_desc(&p.out_map);
        if (p.k2_blocks) {
            tma_prefetch_desc(&p.a2_map);
            tma_prefetch_desc(&p.b2h_map);
        }
        for (int s = 0; s < C::kStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full[a], 1);
            mbar_init(&tmem_empty[a], 8);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_2sm<C::kTmemCols>(tmem_base_slot);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                             // barriers of both CTAs are initialised before any remote signal
    tc_fence_after();
    pdl_launch_dependents();                        // the next kernel in the stream may start its prologue on SMs we leave
    pdl_wait();                                     // our inputs (and buffers we overwrite) belong to the previous kernel until here
    const uint32_t tmem_base = *tmem_base_slot;

    const int taps = p.ksize * p.ksize;
    const int cblocks = p.Cin / kBlockK;
    const int ksteps = taps * cblocks + p.k2_blocks;
    const int m_groups = p.m_tiles_per_img / (2 * MT);       // a pair tile = 2*MT consecutive M tiles of one image
    const int tiles_per_head = p.imgs_per_head * m_groups * p.n_tiles;
    const int total_groups = p.total_tiles / (2 * MT);

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer (both CTAs)
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = pair; tile < total_groups; tile += n_pairs) {
                const int head = tile / tiles_per_head;
                int r = tile - head * tiles_per_head;
                const int n_t = r % p.n_tiles;
                r /= p.n_tiles;
                const int m_t = ((r % m_groups) * 2 + static_cast<int>(rank)) * MT;
                const int img = r / m_groups;
                const int img_in = p.shared_input ? img : head * p.imgs_per_head + img;
                const int oy0 = m_t * p.rows_per_tile;
                const int wrow = head * p.Cout + n_t * N_TILE + static_cast<int>(rank) * (N_TILE / 2);
                for (int tap = 0; tap < taps; ++tap) {
                    const int offy = tap / p.ksize - p.pad;
                    const int offx = tap % p.ksize - p.pad;
                    int map = 0, x0 = offx, y0 = oy0 + offy;
                    if (p.stride == 2) {
                        map = ((offy & 1) << 1) | (offx & 1);
                        x0 = offx >> 1;
                        y0 = oy0 + (offy >> 1);
                    }
                    for (int cb = 0; cb < cblocks; ++cb) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        uint8_t* a_dst = tiles + stage * C::kStageBytes;
                        if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * C::kStageBytes);
#pragma unroll
                        for (int m = 0; m < MT; ++m)
                            tma_load_4d_2sm(a_dst + m * kABytes, &p.a_map[map], &full_bar[stage], cb * kBlockK, x0,
                                            y0 + m * p.rows_per_tile, img_in);
                        tma_load_2d_2sm(a_dst + MT * kABytes, &p.bh_map, &full_bar[stage], tap * p.Cin + cb * kBlockK, wrow);
                        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
                    }
                }
                for (int cb = 0; cb < p.k2_blocks; ++cb) {       // fused downsample branch
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* a_dst = tiles + stage * C::kStageBytes;
                    if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * C::kStageBytes);
#pragma unroll
                    for (int m = 0; m < MT; ++m)
                        tma_load_4d_2sm(a_dst + m * kABytes, &p.a2_map, &full_bar[stage], cb * kBlockK, 0,
                                        oy0 + m * p.rows_per_tile, img_in);
                    tma_load_2d_2sm(a_dst + MT * kABytes, &p.b2h_map, &full_bar[stage], cb * kBlockK, wrow);
                    if (++stage == C::kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ UMMA issuer (leader CTA only)
        if (lane == 0 && rank == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(256, N_TILE);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = pair; tile < total_groups; tile += n_pairs, ++it) {
                const int acc = it & 1;
                mbar_wait(&tmem_empty[acc], ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * C::kAccCols;
                for (int ks = 0; ks < ksteps; ++ks) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(tiles + stage * C::kStageBytes);
                    const uint64_t bdesc = umma_desc_sw128(a_addr + MT * kABytes);
#pragma unroll
                    for (int m = 0; m < MT; ++m) {
                        const uint64_t adesc = umma_desc_sw128(a_addr + m * kABytes);
#pragma unroll
                        for (int k = 0; k < kBlockK / 16; ++k)
                            umma_bf16_2sm(d_tmem + m * N_TILE, adesc + 2 * k, bdesc + 2 * k, idesc, (ks | k) != 0 ? 1u : 0u);
                    }
                    umma_commit_2sm(&empty_bar[stage]);
                    if (++stage == C::kStages) { stage = 0; phase ^= 1; }
                }
                umma_commit_2sm(&tmem_full[acc]);
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue (warps 2..5 of both CTAs)
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        uint8_t* my_out = out_sm + quarter * C::kOutBufs * 4096;
        int it = 0;
        uint32_t nstore = 0;
        for (int tile = pair; tile < total_groups; tile += n_pairs, ++it) {
            const int head = tile / tiles_per_head;
            int r = tile - head * tiles_per_head;
            const int n_t = r % p.n_tiles;
            r /= p.n_tiles;
            const int m_t = ((r % m_groups) * 2 + static_cast<int>(rank)) * MT;
            const int img = r / m_groups;
            const int acc = it & 1;
            const int co0 = n_t * N_TILE;
            const float4* bias4 = reinterpret_cast<const float4*>(p.bias + head * p.Cout + co0);

            mbar_wait(&tmem_full[acc], (it >> 1) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int m = 0; m < MT; ++m) {
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
                if (c0 + 64 >= N_TILE && m == MT - 1) {   // accumulators are in registers: hand TMEM back to the leader's MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_leader(&tmem_empty[acc]);
                }
                if (lane == 0) tma_store_wait_read<C::kOutBufs - 1>();
                __syncwarp();
                uint8_t* stage = my_out + (nstore % C::kOutBufs) * 4096;
#pragma unroll
                for (int ch = 0; ch < 8; ++ch) {
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

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                             // the peer may still signal our barriers / read our smem until here
    if (warp == 1) tmem_dealloc_2sm<C::kTmemCols>(tmem_base);
}

template <int N_TILE, int MT>
cudaError_t launch2_t(const ConvLaunch& p, int num_sms, cudaStream_t stream) {
    using C = Cfg2<N_TILE, MT>;
    cudaError_t e = ensure_dynamic_smem<conv_umma2_kernel<N_TILE, MT>>(C::kSmemBytes);
    if (e != cudaSuccess) {
        fprintf(stderr, "conv_umma2<%d>: cudaFuncSetAttribute(%d B) -> %s\n", N_TILE, C::kSmemBytes, cudaGetErrorString(e));
        return e;
    }
    const int groups = p.total_tiles / (2 * MT);
    int pairs = num_sms / 2;
    if (groups < pairs) pairs = groups;
    e = launch_pdl(conv_umma2_kernel<N_TILE, MT>, dim3(2 * pairs), dim3(kThreads2), C::kSmemBytes, stream, p);
    if (e != cudaSuccess)
        fprintf(stderr, "conv_umma2<%d>: launch grid %d smem %d -> %s\n", N_TILE, 2 * pairs, C::kSmemBytes, cudaGetErrorString(e));
    return e;
}

}  // namespace

cudaError_t conv_umma2_launch(const ConvLaunch& p, int num_sms, cudaStream_t stream) {
    if (p.m_tiles_per_img % 2 != 0) return cudaErrorInvalidValue;
    switch (p.n_tile) {
        case 128: return (p.m_tiles_per_img % 4 == 0) ? launch2_t<128, 2>(p, num_sms, stream) : launch2_t<128, 1>(p, num_sms, stream);
        case 256: return launch2_t<256, 1>(p, num_sms, stream);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace sad
