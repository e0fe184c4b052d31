// kernels_demod_warp.cu — A1/A3/A4 for sm_100a, WARP-PER-STREAM variant (demod_warp_core.cuh).
//
// A stream is a strictly serial recurrence, so with S streams the throughput is S * 40 samples per
// symbol latency.  For small banks (BASELINE configs[1]: 1,024 streams on 148 SMs) that latency is
// all that matters, and this kernel minimises it: one warp owns one stream, the 60-sample window is
// split into 12 five-sample Horner segments per tone (24 lanes + one lane per tone for slot 60),
// the gate sums are formed by three shuffle-down steps, the uniform loop arithmetic (soft decision,
// TED, timing loop) runs redundantly on all lanes, the AFC atan2 runs on the on-time gate lanes.
//
// Samples reach shared memory through an 8-slot ring of 64-sample (256-byte) TMA bulk copies
// (cp.async.bulk + mbarrier complete_tx) issued by lane 0 six slots (~9 symbols) ahead of the window,
// so HBM is read exactly once in 256-byte bursts and its latency is hidden even with one resident
// warp.  Lanes address the ring modulo its size, so no mirror copy is needed.
#include <cuda_runtime.h>
#include <cstdint>

#include "demod_warp_core.cuh"
#include "opvd_kernels.cuh"

namespace opvd {

namespace {

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

constexpr int kSlotShift = 6;
constexpr int kSlotSamples = 1 << kSlotShift;  // 64 samples = 256 B per bulk copy
constexpr int kSlotBytes = kSlotSamples * 4;
constexpr int kNumSlots = 8;
constexpr int kRingWords = kNumSlots * kSlotSamples;  // 512 words = 2 KB per stream
constexpr int kRingMask = kRingWords - 1;
constexpr unsigned kFull = 0xffffffffu;
constexpr int kWarpsPerCta = 1;  // one stream per CTA spreads small banks evenly over the SMs

__device__ __forceinline__ cplx shfl_down_c(cplx v, int d) {
    return {__shfl_down_sync(kFull, v.r, d), __shfl_down_sync(kFull, v.i, d)};
}
__device__ __forceinline__ cplx shfl_c(cplx v, int src) {
    return {__shfl_sync(kFull, v.r, src), __shfl_sync(kFull, v.i, src)};
}

}  // namespace

__global__ void __launch_bounds__(32 * kWarpsPerCta)
demod_warp_kernel(StreamBuffers sb, SoftBuffers so, DemodState* __restrict__ dstate, int n_streams, int mode,
                  int final_flag, double afc_alpha, unsigned long long* __restrict__ counters) {
    __shared__ __align__(128) uint32_t ring_all[kWarpsPerCta][kRingWords];
    __shared__ __align__(8) unsigned long long mbar_all[kWarpsPerCta][kNumSlots];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int stream = blockIdx.x * kWarpsPerCta + wid;
    if (stream >= n_streams) return;  // warp-uniform

    const uint32_t* ring = ring_all[wid];
    const uint32_t ring_s = smem_u32(ring_all[wid]);
    const uint32_t mbar_s = smem_u32(mbar_all[wid]);
    if (lane == 0) {
#pragma unroll
        for (int p = 0; p < kNumSlots; ++p) mbar_init(mbar_s + 8 * p, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();

    DemodState st = dstate[stream];
    const long long avail = sb.avail[stream];
    const long long row0 = sb.row_base;
    const uint32_t* row = sb.iq + (long long)stream * sb.stride;  // row[r] holds absolute sample row0 + r
    const int stride_i = (int)sb.stride;
    const int avail_rel = (int)(avail - row0);
    double* soft_row = so.soft + (long long)stream * so.stride - so.base;

    WarpLane wl;
    warp_lane_init(wl, lane);
    warp_lane_lo(wl, st.freq_offset);
    wl.prev = wl.tone ? st.p2 : st.p1;
    const int pc = wl.p > 12 ? 12 : wl.p;
    const int lane_slot = 5 * pc;
    const int tone_base = lane & 16;

    double freq_offset = st.freq_offset, pos = st.pos, timing_freq = st.timing_freq;
    double ph_own = wl.tone ? st.ph2 : st.ph1;  // this lane's tone's absolute LO phase

    const long long n_sym0 = st.n_sym, origin0 = st.origin;
    double call_len_d = (double)st.call_len;  // 0 => the slow path opens the next call
    int origin_rel = (int)(st.origin - row0);
    int a_first = -1, issued = 0, ready = 0;  // slot indices relative to the row (64-sample units)
    long long n_sym = st.n_sym;
    int sym_in_call = st.sym_in_call;

    for (;;) {
        if (!((pos + 40.0) + 10.0 < call_len_d)) {  // :221 fails or no call open: slow path (warp-uniform)
            st.n_sym = n_sym;
            st.sym_in_call = sym_in_call;
            const bool live = demod_schedule(st, pos, mode, avail, final_flag != 0);
            sym_in_call = st.sym_in_call;
            call_len_d = (double)st.call_len;
            origin_rel = (int)(st.origin - row0);
            if (!live) break;
        }
        const int b = __double2int_rz(pos);  // pos >= 0: truncation == floor (:125)
        const double f = pos - (double)b;
        const int w0 = origin_rel + b - kWinLead;  // row index of window slot 0 (>= -10)
        const int a_lo = w0 >> kSlotShift, a_hi = (w0 + (kWin - 1)) >> kSlotShift;
        __syncwarp();  // every lane is done with the ring slots about to be recycled
        if (a_first < 0) {  // first symbol of this launch: prime the ring
            a_first = a_lo < 0 ? 0 : a_lo;
            issued = ready = a_first;
        }
        // prefetch: slot s overwrites the ring position of slot s-8, which must be behind the window
        while (issued <= a_lo + (kNumSlots - 1) && (issued << kSlotShift) < avail_rel) {
            if (lane == 0) {
                const int p = issued & (kNumSlots - 1);
                const int r0 = issued << kSlotShift;
                const int left = stride_i - r0;
                const uint32_t bytes = left >= kSlotSamples ? (uint32_t)kSlotBytes : (uint32_t)(left * 4);
                const uint32_t mb = mbar_s + 8 * p;
                mbar_expect_tx(mb, bytes);
                tma_bulk_g2s(ring_s + p * kSlotBytes, row + r0, bytes, mb);
            }
            ++issued;
        }
        while (ready <= a_hi) {
            if (ready >= a_first) {
                const uint32_t mb = mbar_s + 8 * (ready & (kNumSlots - 1));
                const uint32_t parity = (uint32_t)(((ready - a_first) >> 3) & 1);
                while (!mbar_try_wait(mb, parity)) {}
            }
            ++ready;
        }

        // ---- lane partial sums over this lane's five slots
        uint32_t s5[5];
        const int w_lane = w0 + lane_slot;
#pragma unroll
        for (int r = 0; r < 5; ++r) s5[r] = ring[(w_lane + r) & kRingMask];
        const LanePartial lp = warp_lane_partial(wl, s5);

        // ---- gate sums: 8 consecutive lanes via shuffle-down 1, 2, 4; edge term from lane p+8
        const cplx Fh = shfl_down_c(lp.F, 8);
        cplx acc = lp.W;
        {
            const cplx o = shfl_down_c(acc, 1);
            acc = {acc.r + o.r, acc.i + o.i};
        }
        {
            const cplx o = shfl_down_c(acc, 2);
            acc = {acc.r + o.r, acc.i + o.i};
        }
        {
            const cplx o = shfl_down_c(acc, 4);
            acc = {acc.r + o.r, acc.i + o.i};
        }
        cplx X = warp_lane_gate(wl, f, acc, Fh, lp.F);
        const bool first = sym_in_call == 0;
        if (first) {  // early-gate clamp at the start of a call (:237); rare, warp-uniform branch
            const cplx fix = first_symbol_fix_w([&](int k) { return ring[(w0 + k) & kRingMask]; }, f, wl.z);
            if (wl.p == kWarpGateLaneE) { X.r -= fix.r; X.i -= fix.i; }
        }
        const double nrm = cnorm(X);

        // ---- energies to every lane: one round of independent shuffles
        const double e1 = __shfl_sync(kFull, nrm, kWarpGateLaneO), e2 = __shfl_sync(kFull, nrm, 16 + kWarpGateLaneO);
        const double eE1 = __shfl_sync(kFull, nrm, kWarpGateLaneE), eL1 = __shfl_sync(kFull, nrm, kWarpGateLaneL);
        const double eE2 = __shfl_sync(kFull, nrm, 16 + kWarpGateLaneE), eL2 = __shfl_sync(kFull, nrm, 16 + kWarpGateLaneL);
        bool tone1;
        const double soft = warp_uniform_timing(e1, e2, eE1, eL1, eE2, eL2, timing_freq, pos, tone1);

        // ---- AFC on the on-time gate lanes (:289-307)
        cplx Ou;
        const double pd_own = warp_lane_afc_phase(wl, X, ph_own, first, Ou);
        const cplx z40 = shfl_c(wl.R, tone_base | kWarpLaneZ40);
        wl.prev = cmul(Ou, cconj(z40));  // :309-310, pre-rotated to the next symbol's phase frame
        ph_own = wrap_phase(fma(40.0, wl.inc, ph_own));  // :250-262
        if (!first) {
            const double pd = __shfl_sync(kFull, pd_own, tone1 ? kWarpGateLaneO : 16 + kWarpGateLaneO);
            afc_loop(freq_offset, pd, afc_alpha);
            warp_lane_lo(wl, freq_offset);
        }
        if (lane == 0) soft_row[n_sym] = soft;
        ++n_sym;
        ++sym_in_call;
    }

    // persist the stream's state
    const double ph1 = __shfl_sync(kFull, ph_own, 0), ph2 = __shfl_sync(kFull, ph_own, 16);
    const cplx p1 = shfl_c(wl.prev, kWarpGateLaneO), p2 = shfl_c(wl.prev, 16 + kWarpGateLaneO);
    if (lane == 0) {
        st.freq_offset = freq_offset; st.pos = pos; st.timing_freq = timing_freq;
        st.ph1 = ph1; st.ph2 = ph2; st.p1 = p1; st.p2 = p2;
        dstate[stream] = st;
        unsigned long long dsym = (unsigned long long)(st.n_sym - n_sym0);
        unsigned long long dsmp = (unsigned long long)(st.origin - origin0);
        if (st.flags & kFlagDone) dsmp = (unsigned long long)(avail - origin0);
        if (dsym) atomicAdd(&counters[kCtrSymbols], dsym);
        if (dsmp) atomicAdd(&counters[kCtrSamples], dsmp);
    }
}

cudaError_t launch_demod_warp(const StreamBuffers& sb, const SoftBuffers& so, DemodState* dstate, int n_streams,
                              int mode, int final_flag, double afc_alpha, unsigned long long* counters,
                              cudaStream_t st) {
    const int grid = (n_streams + kWarpsPerCta - 1) / kWarpsPerCta;
    demod_warp_kernel<<<grid, 32 * kWarpsPerCta, 0, st>>>(sb, so, dstate, n_streams, mode, final_flag, afc_alpha,
                                                           counters);
    return cudaGetLastError();
}

}  // namespace opvd
