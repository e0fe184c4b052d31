// tools/warp_place.cu — development aid: on which SM sub-partition (%warpid & 3) does warp w of a CTA land when
// several CTAs of 3 or 4 warps share an SM?  Decides how the channel-bank kernel must rotate its roles.
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
__global__ void k(int* out, int spin) {
    extern __shared__ int sm[];
    unsigned smid, warpid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(warpid));
    if ((threadIdx.x & 31) == 0) {
        int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
        out[(blockIdx.x * nw + w) * 2] = smid;
        out[(blockIdx.x * nw + w) * 2 + 1] = warpid;
    }
    long long t0 = clock64();
    while (clock64() - t0 < spin) sm[threadIdx.x] = threadIdx.x;   // keep every CTA resident while the others start
}
int main() {
    for (int nw = 3; nw <= 4; ++nw) {
        int grid = 148 * 4, threads = 32 * nw, smem = 50 * 1024;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        int* d; cudaMalloc(&d, grid * nw * 2 * sizeof(int));
        k<<<grid, threads, smem>>>(d, 2000000);
        cudaDeviceSynchronize();
        std::vector<int> h(grid * nw * 2);
        cudaMemcpy(h.data(), d, h.size() * sizeof(int), cudaMemcpyDeviceToHost);
        int hist[4][4] = {};
        for (int b = 0; b < grid; ++b) for (int w = 0; w < nw; ++w) hist[w][h[(b * nw + w) * 2 + 1] & 3]++;
        printf("warps/CTA %d: rows = warp index in CTA, cols = warpid&3\n", nw);
        for (int w = 0; w < nw; ++w) printf("  w%d: %4d %4d %4d %4d\n", w, hist[w][0], hist[w][1], hist[w][2], hist[w][3]);
        printf("  SM 0 CTAs:");
        for (int b = 0; b < grid; ++b) if (h[b * nw * 2] == 0) { printf(" [b%d:", b); for (int w = 0; w < nw; ++w) printf(" %d", h[(b * nw + w) * 2 + 1]); printf("]"); }
        printf("\n");
        cudaFree(d);
    }
    return 0;
}
