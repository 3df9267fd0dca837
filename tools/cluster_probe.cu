// Probe: which (cluster, dynamic smem, threads) launch configurations does this driver accept?  (tools/, not product)
#include <cstdio>
#include <cuda_runtime.h>
struct alignas(64) Big { char b[1536]; int x; };
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1) k_static(const __grid_constant__ Big p, int* out) {
    extern __shared__ unsigned char sm[];
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = p.x + (int)sm[0] * 0;
}
__global__ void __launch_bounds__(192, 1) k_dyn(const __grid_constant__ Big p, int* out) {
    extern __shared__ unsigned char sm[];
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = p.x + (int)sm[0] * 0;
}
int main() {
    int* d; cudaMalloc(&d, 4);
    Big p; p.x = 7;
    for (int kb : {48, 100, 150, 200, 220, 225}) {
        int bytes = kb * 1024;
        cudaError_t e1 = cudaFuncSetAttribute(k_static, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        k_static<<<96, 192, bytes>>>(p, d);
        cudaError_t e2 = cudaGetLastError();
        cudaError_t e3 = cudaDeviceSynchronize();
        printf("static cluster dims, smem %3d KB: attr=%s launch=%s sync=%s\n", kb, cudaGetErrorName(e1), cudaGetErrorName(e2), cudaGetErrorName(e3));
        cudaFuncSetAttribute(k_dyn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(96); cfg.blockDim = dim3(192); cfg.dynamicSmemBytes = bytes;
        cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaError_t e4 = cudaLaunchKernelEx(&cfg, k_dyn, p, d);
        cudaError_t e5 = cudaDeviceSynchronize();
        int nc = -1; cudaOccupancyMaxActiveClusters(&nc, k_dyn, &cfg);
        printf("runtime cluster attr,  smem %3d KB: launch=%s sync=%s maxActiveClusters=%d\n", kb, cudaGetErrorName(e4), cudaGetErrorName(e5), nc);
    }
    return 0;
}
