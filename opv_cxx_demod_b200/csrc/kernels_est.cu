// kernels_est.cu — A2: coarse carrier-offset estimate for sm_100a (reference: estimate_offset,
// /root/reference/src/opv-demod.cpp:131-202) from exact block autocorrelations; see est_core.cuh.
#include <cuda_runtime.h>
#include <cstdint>

#include "opvd_kernels.cuh"
#include "est_core.cuh"

namespace opvd {

// ------------------------------------------------------------------------------------------------
// A2: estimate.  One CTA per stream.  The summed block autocorrelation R[l] = sum_blocks sum_i
// x[i+l] conj(x[i]) is an exact integer (< 2^47): products of int16 values and their sums are exact
// in FP64, so the summation order is free and the work can be register-tiled:
//   * 32 blocks of 40 samples per pass are converted ONCE to double2 in shared memory (the first
//     version converted every operand of every product: 3.3 M I2F per stream, XU-pipe bound);
//   * thread (block slot, role r) owns the 4-lag tiles t = r and t = 9 - r (lags 4t..4t+3) and walks
//     i in steps of 4: 4 x-samples and 7 y-samples in registers feed 16 products, i.e. 0.5 shared
//     loads per product instead of 2.  Tile t needs 10 - t steps, so every role does 11 steps.
//   * blocks are zero-padded to 44 samples, so products that reach past the block need no mask.
// Integer accumulation (one IMAD.WIDE s32 x s32 + s64 per product instead of one DFMA) was measured in round 2 and is
// slower on B200: 12.0 ms against 8.4 ms for 18,944 streams.  The 64-bit integer multiply-add issues at a lower rate
// than the FP64 pipe's DFMA.
// Double-buffered staging (one CTA barrier per pass, the next pass fetched into registers before the products of this
// one) was measured too: 8.24 ms against 8.37 ms, at 96 instead of 66 registers and twice the shared memory - not kept.
constexpr int kEstSlots = 32;                 // blocks per pass
constexpr int kEstRoles = 5;
constexpr int kEstThreads = kEstSlots * kEstRoles;   // 160
constexpr int kEstBlockStride = 45;           // double2 per staged block (40 samples + zero pad; odd: bank spreading)

__device__ __forceinline__ void est_tile(const double2* __restrict__ blk, int t, double (&ar)[4], double (&ai)[4]) {
    // lags L0..L0+3, L0 = 4t; i0 = 0, 4, .., 4*(9-t)
    const int L0 = 4 * t;
    double2 y[7];
#pragma unroll
    for (int k = 0; k < 3; ++k) y[k + 4] = blk[L0 + k];
    for (int i0 = 0; i0 <= 4 * (9 - t); i0 += 4) {
#pragma unroll
        for (int k = 0; k < 3; ++k) y[k] = y[k + 4];
#pragma unroll
        for (int k = 3; k < 7; ++k) y[k] = blk[i0 + L0 + k];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const double2 x = blk[i0 + a];
#pragma unroll
            for (int l = 0; l < 4; ++l) {
                const double2 v = y[a + l];
                ar[l] = fma(v.x, x.x, fma(v.y, x.y, ar[l]));
                ai[l] = fma(v.y, x.x, fma(-v.x, x.y, ai[l]));
            }
        }
    }
}

__global__ void __launch_bounds__(kEstThreads) est_kernel(StreamBuffers sb, DemodState* dstate, double* est_out,
                                                          int n_streams, int mode, int final_flag) {
    const int stream = blockIdx.x;
    if (stream >= n_streams) return;
    __shared__ __align__(16) double2 tile[kEstSlots * kEstBlockStride];
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
    for (int i = threadIdx.x; i < kEstSlots * kEstBlockStride; i += kEstThreads) tile[i] = make_double2(0.0, 0.0);
    __syncthreads();
    if (do_est == 0) return;
    if (do_est == 2) {
        if (threadIdx.x == 0) dstate[stream].flags |= kFlagEstDone;
        return;
    }
    // the estimate always runs on samples [0, 40000) before anything has been consumed: in a ring they still sit at
    // their own offsets (the ring is at least one chunk long)
    const uint32_t* row = sb.iq + (long long)stream * sb.stride;
    const int n_blocks = (int)(n_use / kSps);
    // role-major: a warp = one role x 32 block slots, so its lanes run the same tile loops (no divergence)
    // and hit 32 different blocks at the same index (stride 45 double2: conflict-free 128-bit loads)
    const int slot = threadIdx.x % kEstSlots, role = threadIdx.x / kEstSlots;
    double arA[4] = {0, 0, 0, 0}, aiA[4] = {0, 0, 0, 0}, arB[4] = {0, 0, 0, 0}, aiB[4] = {0, 0, 0, 0};

    for (int blk0 = 0; blk0 < n_blocks; blk0 += kEstSlots) {
        const int nb = min(kEstSlots, n_blocks - blk0);
        // stage: coalesced loads, one conversion per sample
        for (int i = threadIdx.x; i < nb * kSps; i += kEstThreads) {
            double a, b;
            unpack_iq(row[(long long)blk0 * kSps + i], a, b);
            tile[(i / kSps) * kEstBlockStride + (i % kSps)] = make_double2(a, b);
        }
        __syncthreads();
        if (slot < nb) {
            const double2* blk = tile + slot * kEstBlockStride;
            est_tile(blk, role, arA, aiA);
            est_tile(blk, 9 - role, arB, aiB);
        }
        __syncthreads();
    }
#pragma unroll
    for (int l = 0; l < 4; ++l) {
        atomicAdd(&Rr[4 * role + l], arA[l]);
        atomicAdd(&Ri[4 * role + l], aiA[l]);
        atomicAdd(&Rr[4 * (9 - role) + l], arB[l]);
        atomicAdd(&Ri[4 * (9 - role) + l], aiB[l]);
    }
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
    prefer_max_shared(est_kernel);
    est_kernel<<<n_streams, kEstThreads, 0, st>>>(sb, dstate, est_out, n_streams, mode, final_flag);
}


}  // namespace opvd
