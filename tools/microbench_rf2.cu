// tools/microbench_rf2.cu — do the pipes of an SM sub-partition share its register-file read bandwidth?  One warp per
// sub-partition runs DFMAs with three register operands (3.0 cycles each alone, microbench_rf.cu); a second warp on the
// same sub-partition runs integer instructions with three register operands (IADD3 / LOP3).  If the two pipes had
// their own operand paths both would keep their solo rates.  Prints one JSON object.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o tools/microbench_rf2 tools/microbench_rf2.cu
#include <cuda_runtime.h>
#include <cstdio>

// MODE 0: every warp DFMA; 1: every warp integer; 2: warps 0-3 DFMA, warps 4-7 integer (blockDim = 256: warp w sits on
// sub-partition w & 3, so each sub-partition gets one of each)
template <int MODE>
__global__ void k(double* out, long long* clk, int iters, double seed) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, warp = threadIdx.x >> 5;
    const bool fp = MODE == 0 || (MODE == 2 && warp < 4);
    long long t0 = clock64();
    if (fp) {
        double a[8], b[8], d[8];
        for (int i = 0; i < 8; ++i) { a[i] = seed + i + tid; b[i] = 1.0 + 1e-9 * (i + tid); d[i] = 1e-9 * (i + 1 + tid); }
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = fma(a[i], b[i], d[i]);
        double s = 0;
        for (int i = 0; i < 8; ++i) s += a[i];
        if (s == 12345.678) out[0] = s;
    } else {
        unsigned a[8], b[8], d[8];
        for (int i = 0; i < 8; ++i) { a[i] = (unsigned)seed + i + tid; b[i] = 0x9E3779B9u * (i + 1 + tid); d[i] = 0x85EBCA6Bu + i * tid; }
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = (a[i] ^ b[i]) + d[i];  // LOP3 + IADD3 (or one of each fused), three register operands
        unsigned s = 0;
        for (int i = 0; i < 8; ++i) s += a[i];
        if (s == 12345u) out[0] = s;
    }
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) clk[blockIdx.x * (blockDim.x >> 5) + warp] = t1 - t0;
}

template <int MODE>
void run(int sms, int block, int iters, double* fp_cyc, double* int_cyc) {
    double* d; long long* c;
    const int warps = block / 32;
    cudaMalloc(&d, 8); cudaMalloc(&c, sizeof(long long) * sms * warps);
    k<MODE><<<sms, block>>>(d, c, iters / 10, 1.0);
    k<MODE><<<sms, block>>>(d, c, iters, 1.0);
    cudaDeviceSynchronize();
    long long* h = new long long[sms * warps];
    cudaMemcpy(h, c, sizeof(long long) * sms * warps, cudaMemcpyDeviceToHost);
    double fs = 0, is = 0; int fn = 0, in = 0;
    for (int b = 0; b < sms; ++b)
        for (int w = 0; w < warps; ++w) {
            const bool fp = MODE == 0 || (MODE == 2 && w < 4);
            if (fp) { fs += h[b * warps + w]; ++fn; } else { is += h[b * warps + w]; ++in; }
        }
    *fp_cyc = fn ? fs / fn / (8.0 * iters) : 0;   // cycles per warp instruction (8 per iteration)
    *int_cyc = in ? is / in / (8.0 * iters) : 0;  // cycles per loop statement (a ^ b) + d
    delete[] h; cudaFree(d); cudaFree(c);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount, iters = 20000;
    double f0, i0, f1, i1, f2, i2;
    run<0>(sms, 128, iters, &f0, &i0);   // one DFMA warp per sub-partition
    run<1>(sms, 128, iters, &f1, &i1);   // one integer warp per sub-partition
    run<2>(sms, 256, iters, &f2, &i2);   // one of each per sub-partition
    printf("{\"gpu\": \"%s\", \"alone\": {\"dfma3_cycles\": %.3f, \"int3_cycles_per_stmt\": %.3f}, "
           "\"sharing_a_subpartition\": {\"dfma3_cycles\": %.3f, \"int3_cycles_per_stmt\": %.3f}}\n", p.name, f0, i1, f2, i2);
    return 0;
}
