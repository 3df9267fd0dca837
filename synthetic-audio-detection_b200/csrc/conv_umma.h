// Host-visible launch description of the tcgen05 implicit-GEMM convolution (conv_umma.cu).
#pragma once
#include <atomic>
#include <cstdlib>
#include <cuda.h>
#include "act.cuh"
#include <cuda_runtime.h>

namespace sad {

struct alignas(64) ConvLaunch {
    CUtensorMap a_map[4];   // input views; stride 1 uses [0] (all four equal), stride 2 uses parity (py*2+px)
    CUtensorMap b_map;      // packed weights [heads*Cout][taps*Cin] bf16, K-major
    CUtensorMap out_map;    // output viewed as {Cout, pixels}: box {64 ch, 32 px}, SWIZZLE_128B (TMA store)
    CUtensorMap out32_map;  // output as {Cout, pixels} with box {32 ch, 32 px}, no swizzle (transposed-product epilogue)
    CUtensorMap res_map;    // residual viewed as {Cout, pixels}: box {64 ch, 128 px} (TMA load); valid iff residual
    CUtensorMap res32_map;  // residual as {Cout, pixels} with box {32 ch, 32 px}, no swizzle (transposed-product epilogue)
    CUtensorMap a2_map;     // optional second input (fused 1x1/stride-2 downsample branch): parity-(0,0) view of the block input
    CUtensorMap b2_map;     // its weights [heads*Cout][Cin2] bf16; the extra k2_blocks K-steps accumulate into the same tile
    CUtensorMap bh_map;     // b_map / b2_map with box {64, n_tile/2}: each CTA of a 2-CTA pair loads half of the N rows
    CUtensorMap b2h_map;
    const float* bias;      // [heads*Cout] fp32 (folded BN shift)
    const float* bias2;     // fused BasicBlock (block_rows.cu): bias of conv2 (its weights are b2_map)
    const act_t* residual;   // NHWC [heads*imgs][Ho*Wo][Cout] or nullptr
    act_t* out;     // NHWC [heads*imgs][Ho*Wo][Cout]
    int Cin, Cout;
    int ksize, stride, pad;
    int imgs_per_head;      // B
    int m_tiles_per_img;    // Ho*Wo / 128
    int rows_per_tile;      // 128 / Wo
    int n_tiles;            // Cout / n_tile
    int n_tile;             // 64, 128 or 256
    int total_tiles;        // heads * imgs * m_tiles_per_img * n_tiles
    int relu;
    int shared_input;       // 1: every head reads image `img` (stem); 0: head h reads image h*B+img
    int k2_blocks;          // Cin2 / 64 extra K blocks read through a2_map / b2_map (0 = none)
    int transposed;         // N = 128 layers: multiply as weights x pixels (conv_umma.cu, TR); a residual is added in the epilogue through res32_map
};

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device: remember it per (kernel, device).  The kernel is a
// non-type template parameter so every kernel gets its own flags (kernels with the same signature share one TYPE).
template <auto Kernel>
inline cudaError_t ensure_dynamic_smem(int bytes) {
    // per-device flag; atomics because two contexts may launch from two host threads (setting the attribute twice is harmless)
    static std::atomic<bool> done[64];
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64 || !done[dev].load(std::memory_order_acquire)) {
        e = cudaFuncSetAttribute(Kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) done[dev].store(true, std::memory_order_release);
    }
    return cudaSuccess;
}

// Launch with (optionally) the programmatic-stream-serialization attribute: the kernel may begin while its predecessor
// in the stream drains; every kernel launched this way calls pdl_wait() (ptx.cuh) before its first dependent access.
// SAD_PDL=0 turns the attribute off (A/B switch).
inline bool pdl_enabled() {
    static const bool on = [] {
        const char* e = getenv("SAD_PDL");
        return !e || atoi(e) != 0;
    }();
    return on;
}
template <typename Kernel, typename... Args>
inline cudaError_t launch_pdl(Kernel kernel, dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    if (pdl_enabled()) {
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
    }
    return cudaLaunchKernelEx(&cfg, kernel, args...);
}

int conv_n_tile(int Cout);
cudaError_t conv_umma_launch(const ConvLaunch& p, int num_sms, cudaStream_t stream);
// 2-CTA (cta_group::2) variant: M = 256 per CTA pair, B split across the pair.  Requires an even m_tiles_per_img.
cudaError_t conv_umma2_launch(const ConvLaunch& p, int num_sms, cudaStream_t stream);
// Row-stationary 3x3 conv for the 64-channel 128x128 maps (conv_rows.cu) and the whole layer1 BasicBlock on a CTA pair
// (block_rows.cu: conv1 on the even CTA, conv2 + identity on the odd one, the intermediate goes through peer smem).
cudaError_t conv_rows_launch(const ConvLaunch& p, int heads, int num_sms, cudaStream_t stream);
cudaError_t block_rows_launch(const ConvLaunch& p, int heads, int num_sms, cudaStream_t stream);
// conv_rows on CTA pairs (conv_rows2.cu: cta_group::2, UMMA 256x192x16, aliased accumulator ring); needs bh_map = weights with box {64, 32}.
cudaError_t conv_rows2_launch(const ConvLaunch& p, int heads, int num_sms, cudaStream_t stream);

// Tensor-map construction (api.cu): resolves cuTensorMapEncodeTiled through the runtime so that the
// library does not link against libcuda and still loads on a machine without a driver.
// 4-D NHWC bf16 activation view: dims {C, W, H, N}; element (c,x,y,n) at base + c + x*sx + y*sy + n*sn (elements).
bool encode_act_map(CUtensorMap* m, const void* base, int C, int W, int H, long long N, long long sx, long long sy,
                    long long sn, int box_w, int box_h, char* err, int errlen);
// 2-D pixel-major view of an NHWC tensor: dims {C, pixels}; box {64 ch, box_px}.
bool encode_pix_map(CUtensorMap* m, const void* base, int C, long long pixels, int box_px, char* err, int errlen);
// The same view with box {32 ch, 32 px} and no swizzle (dense 64-byte rows in shared memory).
bool encode_pix_map32(CUtensorMap* m, const void* base, int C, long long pixels, char* err, int errlen);
// 2-D weight view: dims {K, rows}; box {64, box_rows}.
bool encode_weight_map(CUtensorMap* m, const void* base, long long K, long long rows, int box_rows, char* err,
                       int errlen);

}  // namespace sad
