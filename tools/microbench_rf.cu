// tools/microbench_rf.cu — does the register file bound DFMA throughput?  DFMA issue rate per SM sub-partition for
// operand patterns with 1, 2 and 3 operands that the operand-reuse cache cannot serve, with many warps per
// sub-partition and with ONE warp per sub-partition (the bank kernel's situation).  Prints one JSON object.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o tools/microbench_rf tools/microbench_rf.cu
#include <cuda_runtime.h>
#include <cstdio>

// P = number of "fresh" 64-bit register operands per DFMA (the others are the same register in the same slot as in
// the previous DFMA, which the reuse cache serves).  8 independent chains, so latency is covered by one warp.
template <int P>
__global__ void k(double* out, int iters, double seed) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    double a[8], b[8], d[8];
    for (int i = 0; i < 8; ++i) { a[i] = seed + i + tid; b[i] = 1.0 + 1e-9 * (i + tid); d[i] = 1e-9 * (i + 1 + tid); }
    const double m = b[0], c = d[0];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (P == 1) a[i] = fma(a[i], m, c);
            if (P == 2) a[i] = fma(a[i], b[i], c);
            if (P == 3) a[i] = fma(a[i], b[i], d[i]);
            if (P == 4) {  // Horner step of the bank kernel: g = g*z + s with z = (m, c) shared by the chains
                // a[i], b[i] = g.r, g.i;  d[i], d[(i+1)&7] stand in for the sample
                const double nr = fma(a[i], m, fma(-b[i], c, d[i]));
                const double ni = fma(a[i], c, fma(b[i], m, d[(i + 1) & 7]));
                a[i] = nr; b[i] = ni;
            }
        }
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) s += a[i] + b[i] + d[i];
    if (s == 12345.678) out[0] = s;
}

template <int P>
double run(int grid, int block, int iters, double ops) {
    double* d; cudaMalloc(&d, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<P><<<grid, block>>>(d, iters / 10, 1.0);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0); k<P><<<grid, block>>>(d, iters, 1.0); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    cudaFree(d);
    return (double)grid * block * iters * ops / (best * 1e-3);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int sms = p.multiProcessorCount, iters = 20000;
    const double per_smsp_clk = 1.0 / ((double)sms * 4 * clk * 1e3 * 32);  // warp-instructions per cycle and sub-partition
    printf("{\"gpu\": \"%s\", \"clock_khz\": %d", p.name, clk);
    // many warps (16 per sub-partition) / one warp per sub-partition
    printf(", \"many_warps_cycles_per_dfma\": {\"fresh1\": %.3f, \"fresh2\": %.3f, \"fresh3\": %.3f, \"horner\": %.3f}",
           1.0 / (run<1>(sms * 8, 256, iters, 8) * per_smsp_clk), 1.0 / (run<2>(sms * 8, 256, iters, 8) * per_smsp_clk),
           1.0 / (run<3>(sms * 8, 256, iters, 8) * per_smsp_clk), 1.0 / (run<4>(sms * 8, 256, iters, 32) * per_smsp_clk));
    printf(", \"one_warp_cycles_per_dfma\": {\"fresh1\": %.3f, \"fresh2\": %.3f, \"fresh3\": %.3f, \"horner\": %.3f}",
           1.0 / (run<1>(sms, 128, iters, 8) * per_smsp_clk), 1.0 / (run<2>(sms, 128, iters, 8) * per_smsp_clk),
           1.0 / (run<3>(sms, 128, iters, 8) * per_smsp_clk), 1.0 / (run<4>(sms, 128, iters, 32) * per_smsp_clk));
    printf(", \"two_warps_cycles_per_dfma\": {\"fresh2\": %.3f, \"fresh3\": %.3f, \"horner\": %.3f}}\n",
           1.0 / (run<2>(sms, 256, iters, 8) * per_smsp_clk), 1.0 / (run<3>(sms, 256, iters, 8) * per_smsp_clk),
           1.0 / (run<4>(sms, 256, iters, 32) * per_smsp_clk));
    return 0;
}
