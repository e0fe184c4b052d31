// tools/microbench.cu — measured issue peaks of the pipes the opv-demod kernels are bound by
// (MEASURED_PEAKS.json only carries HBM copy and bf16 GEMM).  Prints one JSON object.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstring>

template <int OP>
__global__ void k(unsigned long long* out, int iters, double seed) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (OP == 0) {  // DFMA
        double a[8];
        for (int i = 0; i < 8; ++i) a[i] = seed + i + tid;
        const double m = 1.0000001, c = 1e-9;
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = fma(a[i], m, c);
        double s = 0; for (int i = 0; i < 8; ++i) s += a[i];
        if (s == 12345.678) out[0] = 1;
    } else if (OP == 1) {  // FFMA
        float a[8];
        for (int i = 0; i < 8; ++i) a[i] = (float)seed + i + tid;
        const float m = 1.0000001f, c = 1e-9f;
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], m, c);
        float s = 0; for (int i = 0; i < 8; ++i) s += a[i];
        if (s == 12345.678f) out[0] = 1;
    } else if (OP == 2) {  // IADD3 / LOP (alu pipe)
        unsigned a[8];
        for (int i = 0; i < 8; ++i) a[i] = (unsigned)seed + i + tid;
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = (a[i] ^ 0x9E3779B9u) + (unsigned)it;
        unsigned s = 0; for (int i = 0; i < 8; ++i) s += a[i];
        if (s == 12345u) out[0] = 1;
    } else if (OP == 3) {  // DPX __vibmin_s16x2 + add (Viterbi ACS core)
        unsigned a[8];
        for (int i = 0; i < 8; ++i) a[i] = (unsigned)seed + i + tid;
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int i = 0; i < 8; ++i) { bool ph, pl; a[i] = __vibmin_s16x2(a[i] + 0x00030004u, a[(i + 1) & 7], &ph, &pl) + (ph ? 1u : 0u); }
        unsigned s = 0; for (int i = 0; i < 8; ++i) s += a[i];
        if (s == 12345u) out[0] = 1;
    } else if (OP == 4) {  // SHFL
        unsigned a[4];
        for (int i = 0; i < 4; ++i) a[i] = (unsigned)seed + i + tid;
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = __shfl_sync(0xffffffffu, a[i], (tid + it) & 31);
        unsigned s = 0; for (int i = 0; i < 4; ++i) s += a[i];
        if (s == 12345u) out[0] = 1;
    } else if (OP == 6) {  // I2F.F64.S16 (sample unpack) interleaved with a DADD so it cannot be hoisted
        double a[8];
        unsigned w = (unsigned)seed * 2654435761u + tid;
        for (int i = 0; i < 8; ++i) a[i] = 0.0;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { a[i] += (double)(short)((w >> ((i & 1) * 16)) & 0xFFFFu); w = w * 1664525u + 1013904223u; }
        }
        double s = 0; for (int i = 0; i < 8; ++i) s += a[i];
        if (s == 12345.678) out[0] = 1;
    } else if (OP == 7) {  // magic-number int16 -> double (PRMT/LOP + DADD), same harness as OP 6
        double a[8];
        unsigned w = (unsigned)seed * 2654435761u + tid;
        for (int i = 0; i < 8; ++i) a[i] = 0.0;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const unsigned h = ((w >> ((i & 1) * 16)) & 0xFFFFu) ^ 0x8000u;
                a[i] += __hiloint2double(0x43300000, (int)h) - 4503599627403264.0;
                w = w * 1664525u + 1013904223u;
            }
        }
        double s = 0; for (int i = 0; i < 8; ++i) s += a[i];
        if (s == 12345.678) out[0] = 1;
    } else if (OP == 8) {  // accuracy of the MUFU.RCP64H seed: max |1 - b*r| over a sweep, stored as bits
        double worst = 0.0;
        for (int it = 0; it < iters; ++it) {
            const double b = 1.0 + (double)((tid * 7919u + it * 104729u) & 0xFFFFFu) / 1048576.0 + 1e-9 * it;
            double r;
            asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
            const double e = fabs(fma(-b, r, 1.0));
            worst = e > worst ? e : worst;
        }
        atomicMax(out, (unsigned long long)__double_as_longlong(worst));
    } else if (OP == 5) {  // dependent DFMA chain: latency
        double a = seed + tid;
        for (int it = 0; it < iters * 8; ++it) a = fma(a, 1.0000001, 1e-9);
        if (a == 12345.678) out[0] = 1;
    }
}

template <int OP>
double run(int grid, int block, int iters, double ops_per_iter_thread) {
    unsigned long long* d; cudaMalloc(&d, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<grid, block>>>(d, iters / 10, 1.0);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0); k<OP><<<grid, block>>>(d, iters, 1.0); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    cudaFree(d);
    return (double)grid * block * iters * ops_per_iter_thread / (best * 1e-3);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount, grid = sms * 8, block = 256, iters = 20000;
    double dfma = run<0>(grid, block, iters, 8), ffma = run<1>(grid, block, iters, 8), alu = run<2>(grid, block, iters, 16),
           dpx = run<3>(grid, block, iters, 8), shfl = run<4>(grid, block, iters, 4), i2f = run<6>(grid, block, iters, 8);
    // latency: one warp per SM
    double chain = run<5>(sms, 32, 4000, 8);  // dependent ops/s over sms*32 threads
    double magic = run<7>(grid, block, iters, 8);
    double rcp_err;
    {
        unsigned long long* d; cudaMalloc(&d, 8); cudaMemset(d, 0, 8);
        k<8><<<grid, block>>>(d, 2000, 1.0);
        unsigned long long bits; cudaMemcpy(&bits, d, 8, cudaMemcpyDeviceToHost); cudaFree(d);
        memcpy(&rcp_err, &bits, 8);
    }
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz_max\": %d, \"dfma_per_s\": %.4g, \"fp64_tflops\": %.3f, "
           "\"ffma_per_s\": %.4g, \"fp32_tflops\": %.3f, \"alu_ops_per_s\": %.4g, \"dpx_vibmin_add_per_s\": %.4g, "
           "\"shfl_per_s\": %.4g, \"i2f_f64_s16_plus_dadd_per_s\": %.4g, \"magic_cvt_plus_dadd_per_s\": %.4g, "
           "\"rcp64h_seed_max_rel_err\": %.3g, \"dfma_dependent_ns\": %.3f}\n",
           p.name, sms, clk, dfma, 2 * dfma / 1e12, ffma, 2 * ffma / 1e12, alu, dpx, shfl, i2f, magic, rcp_err,
           1e9 / (chain / (sms * 32.0)));
    return 0;
}
