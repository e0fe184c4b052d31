// kernels_demod.cu — front-end kernels for sm_100a:
//   est_kernel    A2  carrier-offset estimate from exact block autocorrelations (est_core.cuh)
//   demod_kernel  A1/A3/A4  int16 I/Q unpack, dual-tone correlator, early-late STR, AFC,
//                 batch / streaming call schedule (demod_core.cuh)
//
// demod_kernel layout: one stream per lane, one warp per CTA.  A stream is a strictly serial
// recurrence (symbol n+1's window and LO step depend on symbol n), so the parallel axis is
// streams.  Each lane owns a 4-slot ring of 64-sample slots in shared memory (+ a mirror of
// slot 0 so a 61-sample window is always contiguous) that is filled by per-lane TMA bulk copies
// (cp.async.bulk + mbarrier complete_tx) two slots ahead of the symbol being demodulated: HBM is
// read in 256-byte contiguous bursts per stream, exactly once, and the LSU never touches global
// memory on the sample path.  The arithmetic is FP64 (the reference is FP64 and frame parity is
// decided at a 3-bit quantiser), restructured to ~24 FP64 ops per sample (demod_core.cuh).
#include <cuda_runtime.h>
#include <cstdint>

#include "opvd_kernels.cuh"
#include "est_core.cuh"

namespace opvd {

// ------------------------------------------------------------------------------------------------
// PTX helpers (mbarrier + TMA 1-D bulk copy)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t mbar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(mbar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(mbar)
        : "memory");
}

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

// ------------------------------------------------------------------------------------------------
// A1/A3/A4: demodulator.
constexpr int kSlotSamples = 64;                       // 256 B per TMA bulk copy
constexpr int kSlotBytes = kSlotSamples * 4;
constexpr int kNumSlots = 4;
constexpr int kRingWords = (kNumSlots + 1) * kSlotSamples;  // + mirror of ring slot 0
constexpr int kRingStrideBytes = kRingWords * 4 + 16;  // 1296 B: odd multiple of 16 B (bank spreading)
constexpr int kDemodWarpsPerCta = 1;

__global__ void __launch_bounds__(32 * kDemodWarpsPerCta)
demod_kernel(StreamBuffers sb, SoftBuffers so, DemodState* __restrict__ dstate, int n_streams, int mode,
             int final_flag, double afc_alpha, unsigned long long* __restrict__ counters) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x;
    const int stream = blockIdx.x * 32 + lane;
    const bool active = stream < n_streams;

    unsigned char* ring_b = smem + (size_t)lane * kRingStrideBytes;
    const uint32_t* ring = reinterpret_cast<const uint32_t*>(ring_b);
    const uint32_t ring_s = smem_u32(ring_b);
    const uint32_t mbar_s = smem_u32(smem + 32 * kRingStrideBytes + lane * (kNumSlots * 8));

#pragma unroll
    for (int p = 0; p < kNumSlots; ++p) mbar_init(mbar_s + 8 * p, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (!active) return;

    DemodState st = dstate[stream];
    const long long avail = sb.avail[stream];
    const long long row0 = sb.row_base;
    const uint32_t* row = sb.iq + (long long)stream * sb.stride - row0;  // row[g] valid for row0 <= g < row0 + stride
    const long long row_end = row0 + sb.stride;
    double* soft_row = so.soft + (long long)stream * so.stride - so.base;

    DemodRegs r;
    regs_from_state(r, st);
    const long long n_sym0 = st.n_sym;
    const long long origin0 = st.origin;

    long long a_first = -1, next_issue = 0, ready_upto = 0;  // absolute slot indices (64-sample units)

    while (demod_schedule(st, r.pos, mode, avail, final_flag != 0)) {
        const long long b = (long long)r.pos;
        const double f = r.pos - (double)b;
        const long long g0 = st.origin + b - kWinLead;  // global index of window slot 0 (may be < 0)
        const long long a_lo = g0 >> 6, a_hi = (g0 + (kWin - 1)) >> 6;
        if (a_first < 0) {
            a_first = a_lo < 0 ? 0 : a_lo;
            next_issue = ready_upto = a_first;
        }
        // prefetch: slot a may overwrite ring position of slot a-4, which must be behind the window
        while (next_issue <= a_lo + (kNumSlots - 1) && (next_issue << 6) < avail) {
            const int p = (int)(next_issue & (kNumSlots - 1));
            const long long g = next_issue << 6;
            long long left = row_end - g;
            const uint32_t bytes = left >= kSlotSamples ? (uint32_t)kSlotBytes : (uint32_t)(left * 4);
            const uint32_t mb = mbar_s + 8 * p;
            mbar_expect_tx(mb, p == 0 ? 2 * bytes : bytes);
            tma_bulk_g2s(ring_s + p * kSlotBytes, row + g, bytes, mb);
            if (p == 0) tma_bulk_g2s(ring_s + kNumSlots * kSlotBytes, row + g, bytes, mb);
            ++next_issue;
        }
        while (ready_upto <= a_hi) {
            if (ready_upto >= a_first) {
                const uint32_t mb = mbar_s + 8 * (int)(ready_upto & (kNumSlots - 1));
                const uint32_t parity = (uint32_t)(((ready_upto - a_first) >> 2) & 1);
                while (!mbar_try_wait(mb, parity)) {}
            }
            ++ready_upto;
        }
        const uint32_t* win = ring + ((int)(a_lo & (kNumSlots - 1)) * kSlotSamples + (int)(g0 - (a_lo << 6)));
        const double soft = demod_symbol(r, win, f, st.sym_in_call == 0, afc_alpha);
        soft_row[st.n_sym] = soft;
        st.n_sym++;
        st.sym_in_call++;
    }
    regs_to_state(r, st);
    dstate[stream] = st;

    // counters: symbols produced and samples consumed in this launch
    unsigned long long dsym = (unsigned long long)(st.n_sym - n_sym0);
    unsigned long long dsmp = (unsigned long long)(st.origin - origin0);
    if (st.flags & kFlagDone) dsmp = (unsigned long long)(avail - origin0);
    if (dsym) atomicAdd(&counters[kCtrSymbols], dsym);
    if (dsmp) atomicAdd(&counters[kCtrSamples], dsmp);
}

cudaError_t launch_demod(const StreamBuffers& sb, const SoftBuffers& so, DemodState* dstate, int n_streams,
                         int mode, int final_flag, double afc_alpha, int lanes_per_stream,
                         unsigned long long* counters, cudaStream_t st) {
    (void)lanes_per_stream;
    if (n_streams <= 0) return cudaSuccess;
    const size_t smem = 32 * kRingStrideBytes + 32 * kNumSlots * 8;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(demod_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_done = true;
    }
    const int grid = (n_streams + 31) / 32;
    demod_kernel<<<grid, 32, smem, st>>>(sb, so, dstate, n_streams, mode, final_flag, afc_alpha, counters);
    return cudaGetLastError();
}

}  // namespace opvd
