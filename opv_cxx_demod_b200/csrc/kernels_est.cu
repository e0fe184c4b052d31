// kernels_est.cu — A2: coarse carrier-offset estimate for sm_100a (reference: estimate_offset,
// /root/reference/src/opv-demod.cpp:131-202) from exact block autocorrelations; see est_core.cuh.
#include <cuda_runtime.h>
#include <cstdint>

#include "opvd_kernels.cuh"
#include "est_core.cuh"

namespace opvd {

// ------------------------------------------------------------------------------------------------
// A2: estimate.  One CTA per stream, 160 threads = 8 sample-block slots x 20 lag pairs.
// Lag pair p handles lags p and 39-p (41 products per 40-sample block in total), so the work per
// thread is uniform.  Products of int16 and their sums (< 2^47) are exact in FP64, so the
// reduction order is irrelevant and shared-memory atomics can be used.
constexpr int kEstThreads = 160;
constexpr int kEstTile = 25;  // 40-sample blocks staged per pass (1000 samples = 4 KB)

__global__ void __launch_bounds__(kEstThreads) est_kernel(StreamBuffers sb, DemodState* dstate, double* est_out,
                                                          int n_streams, int mode, int final_flag) {
    const int stream = blockIdx.x;
    if (stream >= n_streams) return;
    __shared__ uint32_t tile[kEstTile * kSps];
    __shared__ double Rr[kEstLags], Ri[kEstLags];
    __shared__ double energy[128];
    __shared__ int do_est;
    __shared__ long long n_use;

    if (threadIdx.x == 0) {
        const DemodState& s = dstate[stream];
        const long long avail = sb.avail[stream];
        int e = 0;
        long long n = 0;
        if (!(s.flags & kFlagEstDone)) {
            if (mode == kModeBatch) {
                if (final_flag) { e = 1; n = avail; }                 // :1166 whole capture (first 40,000 used)
            } else if (avail >= kChunkSamples) { e = 1; n = kChunkSamples; }  // :1030-1033 first full chunk
            else if (final_flag) { e = 2; }                           // short stream: never estimated
        }
        do_est = e;
        n_use = n < kEstSamples ? n : kEstSamples;
    }
    if (threadIdx.x < kEstLags) { Rr[threadIdx.x] = 0.0; Ri[threadIdx.x] = 0.0; }
    __syncthreads();
    if (do_est == 0) return;
    if (do_est == 2) {
        if (threadIdx.x == 0) dstate[stream].flags |= kFlagEstDone;
        return;
    }
    const long long row0 = sb.row_base;  // estimate always runs on samples [0, 40000)
    const uint32_t* row = sb.iq + (long long)stream * sb.stride - row0;
    const int n_blocks = (int)(n_use / kSps);
    const int slot = threadIdx.x / 20, pair = threadIdx.x % 20;
    const int lagA = pair, lagB = kSps - 1 - pair;  // lags 0..19 and 39..20: 41 products per block for every pair
    double arA = 0, aiA = 0, arB = 0, aiB = 0;

    for (int blk0 = 0; blk0 < n_blocks; blk0 += kEstTile) {
        const int nb = min(kEstTile, n_blocks - blk0);
        for (int i = threadIdx.x; i < nb * kSps; i += kEstThreads) tile[i] = row[(long long)blk0 * kSps + i];
        __syncthreads();
        for (int b = slot; b < nb; b += 8) {
            const uint32_t* s = tile + b * kSps;
            // lag A: i' = 0 .. 39-lagA ; lag B: i' = 0 .. 39-lagB
            for (int i = 0; i + lagA < kSps; ++i) {
                double a, bq, a2, b2;
                unpack_iq(s[i], a, bq);
                unpack_iq(s[i + lagA], a2, b2);
                arA = fma(a2, a, fma(b2, bq, arA));
                aiA = fma(b2, a, fma(-a2, bq, aiA));
            }
            for (int i = 0; i + lagB < kSps; ++i) {
                double a, bq, a2, b2;
                unpack_iq(s[i], a, bq);
                unpack_iq(s[i + lagB], a2, b2);
                arB = fma(a2, a, fma(b2, bq, arB));
                aiB = fma(b2, a, fma(-a2, bq, aiB));
            }
        }
        __syncthreads();
    }
    atomicAdd(&Rr[lagA], arA);
    atomicAdd(&Ri[lagA], aiA);
    if (lagB < kSps) { atomicAdd(&Rr[lagB], arB); atomicAdd(&Ri[lagB], aiB); }
    __syncthreads();

    // coarse grid: 121 candidates in parallel, then the reference's sequential strict-'>' scan
    if (threadIdx.x < 121) energy[threadIdx.x] = est_energy(Rr, Ri, -1500.0 + 25.0 * threadIdx.x);
    __syncthreads();
    __shared__ double best_offset_s, best_energy_s;
    if (threadIdx.x == 0) {
        double best_offset = 0, best_energy = 0;
        for (int c = 0; c < 121; ++c)
            if (energy[c] > best_energy) { best_energy = energy[c]; best_offset = -1500.0 + 25.0 * c; }
        best_offset_s = best_offset;
        best_energy_s = best_energy;
    }
    __syncthreads();
    if (threadIdx.x < 13) energy[threadIdx.x] = est_energy(Rr, Ri, best_offset_s - 30.0 + 5.0 * threadIdx.x);
    __syncthreads();
    if (threadIdx.x == 0) {
        double fine_best = best_offset_s, best_energy = best_energy_s;
        for (int c = 0; c < 13; ++c)
            if (energy[c] > best_energy) { best_energy = energy[c]; fine_best = best_offset_s - 30.0 + 5.0 * c; }
        dstate[stream].freq_offset = fine_best;
        if (est_out) est_out[stream] = fine_best;
        dstate[stream].flags |= kFlagEstDone;
    }
}

void launch_estimate(const StreamBuffers& sb, DemodState* dstate, double* est_out, int n_streams, int mode,
                     int final_flag, cudaStream_t st) {
    if (n_streams <= 0) return;
    est_kernel<<<n_streams, kEstThreads, 0, st>>>(sb, dstate, est_out, n_streams, mode, final_flag);
}


}  // namespace opvd
