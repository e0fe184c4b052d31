// kernels_demod_warp.cu — A1/A3/A4 for sm_100a, WARP-PER-STREAM variant (demod_warp_core.cuh).
//
// A stream is a strictly serial recurrence, so with S streams the throughput is S * 40 samples per
// symbol latency.  For small banks (BASELINE configs[1]: 1,024 streams on 148 SMs) that latency is
// all that matters, and this kernel minimises it: one warp owns one stream, the 60-sample window is
// split into 12 five-sample Horner segments per tone (24 lanes + one lane per tone for slot 60),
// the gate sums are formed by three shuffle-down steps, the uniform loop arithmetic (soft decision,
// TED, timing loop) runs redundantly on all lanes, the AFC atan2 runs on the on-time gate lanes.
//
// Samples reach shared memory through an 8-slot ring of 64-sample (256-byte) slots filled by 16-byte
// cp.async copies (lanes 0-15, one fully coalesced 256-byte burst per slot) issued seven slots (~11
// symbols) ahead of the window, so HBM is read exactly once, no register (and no register scoreboard)
// is tied to a load in flight, and its latency is hidden even with one resident warp.  The first
// version used cp.async.bulk + mbarriers issued by lane 0: its bookkeeping (uniform-datapath address
// arithmetic, expect_tx, try_wait, two loops per symbol) was 95 of the 319 instructions and ~400 of the
// ~1,340 cycles of a symbol (profiles/ncu_demod_warp_r01_d.txt and the source-level capture).
#include <cuda_runtime.h>
#include <cstdint>

#include "demod_warp_core.cuh"
#include "opvd_kernels.cuh"

namespace opvd {

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int kSlotShift = 6;
constexpr int kSlotSamples = 1 << kSlotShift;  // 64 samples = 256 B per bulk copy
constexpr int kSlotBytes = kSlotSamples * 4;
constexpr int kNumSlots = 8;
constexpr int kRingSamples = kNumSlots * kSlotSamples;     // 512 samples = 2 KB per stream
constexpr int kRingWords = kRingSamples + kSlotSamples;    // + mirror of ring slot 0: windows never wrap
constexpr int kRingMask = kRingSamples - 1;
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ cplx shfl_down_c(cplx v, int d) {
    return {__shfl_down_sync(kFull, v.r, d), __shfl_down_sync(kFull, v.i, d)};
}
__device__ __forceinline__ cplx shfl_c(cplx v, int src) {
    return {__shfl_sync(kFull, v.r, src), __shfl_sync(kFull, v.i, src)};
}

// Everything one warp carries through the symbol loop of its stream.
struct WarpCtx {
    FastMathTable K;     // polynomial / loop constants pinned in registers
    WarpLane wl;
    cplx RP;             // R * prev (gate lane O): O_n * conj(prev) = X * conj(RP)
    bool prev_zero;      // prev is exactly zero (signed-zero corner of the AFC phase detector)
    double freq_offset, pos, timing_freq, ph_own, afc_alpha;
    const uint32_t* ring;
    uint32_t ring_s;
    const uint32_t* row; // this stream's sample row (linear or ring)
    int row_base_off;    // physical offset of rel 0, ring length (INT_MAX for a linear row), first rel beyond a linear row
    int row_wrap, rel_end;
    int issued_off;      // physical offset of issued_s
    int avail_rel, origin_rel;
    int issued_s;        // samples [.., issued_s) of the row have been requested (multiple of 64); < 0: ring not primed
    int lane, lane_slot, tone_base;
    double* soft_row;    // soft-symbol row of this stream (linear or ring)
    int soft_idx, soft_wrap, n_new;  // row position of the next soft symbol (lane 0 stores), ring length, symbols produced

    __device__ __forceinline__ void set_view(const RowView& v) {
        row = v.row; row_base_off = v.base_off; row_wrap = v.wrap; rel_end = v.rel_end; issued_off = 0;
    }

    // request the next 64-sample slot: lanes 0-15 copy 16 bytes each (slot 0 of the ring also feeds the mirror
    // behind the ring's end, so a window never wraps)
    __device__ __forceinline__ void issue_slot() {
        const int p = (issued_s >> kSlotShift) & (kNumSlots - 1);
        const int off = issued_s + 4 * lane;
        if (lane < 16 && off + 4 <= rel_end) {  // linear rows are a multiple of 4 samples long; a ring (a multiple of 64)
            const uint32_t* src = row + issued_off + 4 * lane;  // never splits a 64-sample slot
            cp_async16(ring_s + p * kSlotBytes + 16 * lane, src);
            if (p == 0) cp_async16(ring_s + kNumSlots * kSlotBytes + 16 * lane, src);
        }
        issued_s += kSlotSamples;
        issued_off += kSlotSamples;
        if (issued_off >= row_wrap) issued_off -= row_wrap;
    }

    // Once per symbol: keep 7 slots requested ahead of the window starting at row index w0 (slot s recycles the
    // ring position of slot s-8, whose samples are all older than w0), and have every copy group older than four
    // symbols landed.  A slot is requested when issued_s <= w0 + 448, i.e. at least 448 - 64 - 105 = 279 samples
    // (6.5 symbols) before the window of the NEXT symbol (it ends at most at w0 + 105) can touch it, so the next
    // window can be loaded mid-symbol without another check.  One slot per symbol is enough: a symbol consumes at
    // most 43 samples.
    __device__ __forceinline__ void ring_maintain(int w0) {
        if (w0 + (kNumSlots - 1) * kSlotSamples >= issued_s && issued_s < avail_rel) {
            __syncwarp();  // every lane is done with the slot about to be recycled
            issue_slot();
            // the absolute LO phase is only read in the signed-zero corner: advance it unwrapped in
            // the symbol loop and wrap it here, every ~1.6 symbols (|ph| stays < 10 rad)
            ph_own = warp_wrap_phase(ph_own, K);
        }
        cp_async_commit();
        cp_async_wait<4>();
        __syncwarp();  // the copies of lanes 0-15 are visible to every lane
    }
    // first symbol of a launch: fill the ring up to seven slots ahead of the window and wait for all of it
    __device__ __forceinline__ void ring_prime(int w0) {
        issued_s = w0 < 0 ? 0 : (w0 & ~(kSlotSamples - 1));
        issued_off = row_base_off + issued_s;
        while (issued_off >= row_wrap) issued_off -= row_wrap;
        while (w0 + (kNumSlots - 1) * kSlotSamples >= issued_s && issued_s < avail_rel) issue_slot();
        cp_async_commit();
        cp_async_wait<0>();
        __syncwarp();
    }
    // this lane's five samples of the window at integer position b, converted: the conversions (XU pipe, 8 cycles per
    // warp instruction, ten per symbol) run where they are issued, in the shadow of the AFC chain of the previous symbol,
    // not at the head of the next one where the Horner chain would wait for them
    __device__ __forceinline__ void load_window(int b, double (&I5)[5], double (&Q5)[5]) const {
        const uint32_t* src = ring + (((origin_rel + b - kWinLead) & kRingMask) + lane_slot);
        uint32_t s5[5];
#pragma unroll
        for (int r = 0; r < 5; ++r) s5[r] = src[r];
#pragma unroll
        for (int r = 0; r < 5; ++r) unpack_iq(s5[r], I5[r], Q5[r]);
    }

    // One symbol at integer position b = trunc(pos), whose five lane samples are already in I5/Q5.
    // Updates pos, then loads the NEXT symbol's samples into I5/Q5 and returns its b, before running the
    // AFC chain, so that the loads and the AFC arithmetic overlap.  FIRST: first symbol of a
    // demodulate() call (early-gate clamp :237, no AFC update :289).
    template <bool FIRST>
    __device__ __forceinline__ int symbol(int b, double (&I5)[5], double (&Q5)[5]) {
        const double f = pos - (double)b;
        // ---- lane partial sums over this lane's five slots
        const LanePartial lp = warp_lane_partial_d(wl, I5, Q5);

        // ---- gate sums: 8 consecutive lanes via shuffle-down 1, 2, 4; edge term from lane p+8
        const cplx Fh = shfl_down_c(lp.F, 8);
        cplx acc = lp.W;
#pragma unroll
        for (int d = 1; d <= 4; d <<= 1) {
            const cplx o = shfl_down_c(acc, d);
            acc = {acc.r + o.r, acc.i + o.i};
        }
        cplx X = warp_lane_gate(wl, f, acc, Fh, lp.F);
        if (FIRST) {
            const uint32_t* win = ring + ((origin_rel + b - kWinLead) & kRingMask);
            const cplx fix = first_symbol_fix_w([&](int k) { return win[k]; }, f, wl.z);
            if (wl.p == kWarpGateLaneE) { X.r -= fix.r; X.i -= fix.i; }
        }
        const double nrm = cnorm(X);

        // ---- energies to every lane: one round of independent shuffles
        const double e1 = __shfl_sync(kFull, nrm, kWarpGateLaneO), e2 = __shfl_sync(kFull, nrm, 16 + kWarpGateLaneO);
        const double eE1 = __shfl_sync(kFull, nrm, kWarpGateLaneE), eL1 = __shfl_sync(kFull, nrm, kWarpGateLaneL);
        const double eE2 = __shfl_sync(kFull, nrm, 16 + kWarpGateLaneE), eL2 = __shfl_sync(kFull, nrm, 16 + kWarpGateLaneL);
        bool tone1;
        const double soft = warp_uniform_timing(e1, e2, eE1, eL1, eE2, eL2, timing_freq, pos, tone1, K);
        if (lane == 0) soft_row[soft_idx] = soft;
        if (++soft_idx == soft_wrap) soft_idx = 0;
        ++n_new;

        // ---- next symbol's samples (its window is already in the ring, see ring_maintain)
        const int b_next = __double2int_rz(pos);  // pos >= 0: truncation == floor (:125)
        load_window(b_next, I5, Q5);

        // ---- AFC on the on-time gate lanes (:289-307)
        const bool x_zero = nrm == 0.0;
        double pd_own = 0.0;
        if (!FIRST) pd_own = warp_lane_afc_phase(wl, X, RP, x_zero || prev_zero, ph_own, K);
        const cplx z50 = shfl_c(wl.R, tone_base | 10);   // R of lane p = 10 is z^50 = z^10 * z^40
        wl.prev = cmul(X, cconj(z50));                    // :309-310, pre-rotated to the next symbol's phase frame
        prev_zero = x_zero;
        ph_own = fma(40.0, wl.inc, ph_own);               // :250-262 (wrapped in ring_maintain)
        if (!FIRST) {
            const double pd = __shfl_sync(kFull, pd_own, tone1 ? kWarpGateLaneO : 16 + kWarpGateLaneO);
            warp_afc_loop(freq_offset, pd, afc_alpha, K);
            warp_lane_lo_fast(wl, freq_offset, K);
        }
        RP = cmul(wl.R, wl.prev);
        return b_next;
    }
};

}  // namespace

// MINB: resident CTAs per SM the register budget is cut for.  8 (188 registers) is what ptxas takes unconstrained;
// 16 (124 registers, still no spill, 12 % more time per symbol) doubles the streams that fit in one wave, from 1,184
// to 2,368 on a B200: banks just beyond one wave of the fast build would otherwise take two.
template <int MINB>
__global__ void __launch_bounds__(32, MINB)
demod_warp_kernel(StreamBuffers sb, SoftBuffers so, DemodState* __restrict__ dstate, int n_streams, int mode,
                  int final_flag, double afc_alpha, unsigned long long* __restrict__ counters) {
    __shared__ __align__(128) uint32_t ring_sm[kRingWords];
    const int lane = threadIdx.x;
    const int stream = blockIdx.x;

    WarpCtx c;
    c.K = load_table_pinned();
    c.lane = lane;
    c.ring = ring_sm;
    c.ring_s = smem_u32(ring_sm);

    DemodState st = dstate[stream];
    const long long avail = sb.avail[stream];
    const RowView view = make_row_view(sb, stream, st.origin);
    c.set_view(view);
    const long long row0 = view.base_abs;  // rel = absolute sample index - row0
    c.avail_rel = (int)(avail - row0);
    c.soft_row = so.soft + (long long)stream * so.stride;
    c.soft_wrap = so.ring ? (int)so.stride : 0x7fffffff;
    c.soft_idx = (int)soft_pos(so, st.n_sym);
    c.n_new = 0;
    c.afc_alpha = afc_alpha;

    warp_lane_init(c.wl, lane, c.K);
    warp_lane_lo(c.wl, st.freq_offset);  // general version: a -o offset may exceed the fast range
    c.wl.prev = c.wl.tone ? st.p2 : st.p1;
    c.prev_zero = c.wl.prev.r == 0.0 && c.wl.prev.i == 0.0;
    c.RP = cmul(c.wl.R, c.wl.prev);
    c.lane_slot = 5 * (c.wl.p > 12 ? 12 : c.wl.p);
    c.tone_base = lane & 16;
    c.freq_offset = st.freq_offset;
    c.pos = st.pos;
    c.timing_freq = st.timing_freq;
    c.ph_own = c.wl.tone ? st.ph2 : st.ph1;  // this lane's tone's absolute LO phase
    c.issued_s = -1;

    const long long n_sym0 = st.n_sym, origin0 = st.origin;
    for (;;) {
        // ---- call boundary (warp-uniform slow path): close / open demodulate() calls (:1012-1113)
        st.n_sym = n_sym0 + c.n_new;
        if (!demod_schedule(st, c.pos, mode, avail, final_flag != 0)) break;
        c.origin_rel = (int)(st.origin - row0);
        asm volatile("" : "+r"(c.origin_rel));  // keep it in a register: ptxas otherwise re-derives it from the constant
                                                // bank (row_base) in every symbol, an exposed 26-cycle LDC
        const int call_len_i = (int)st.call_len;
        const double call_len_d = (double)st.call_len;
        int b = __double2int_rz(c.pos);  // pos >= 0: truncation == floor (:125)
        double I5[5], Q5[5];
        if (c.issued_s < 0) c.ring_prime(c.origin_rel + b - kWinLead);  // first symbol of this launch
        else c.ring_maintain(c.origin_rel + b - kWinLead);
        c.load_window(b, I5, Q5);
        if (st.sym_in_call == 0) {
            b = c.symbol<true>(b, I5, Q5);
            st.sym_in_call = 1;
            if (!((c.pos + 40.0) + 10.0 < call_len_d)) continue;
        }
        // ---- hot loop: while (pos + 50 < N) (:221); the integer test is a conservative shortcut
        for (;;) {
            if (b + 52 >= call_len_i && !((c.pos + 40.0) + 10.0 < call_len_d)) break;
            c.ring_maintain(c.origin_rel + b - kWinLead);
            b = c.symbol<false>(b, I5, Q5);
        }
        st.sym_in_call = 2;  // any non-zero value: the open call has produced symbols
    }

    // persist the stream's state
    c.ph_own = warp_wrap_phase(c.ph_own, c.K);
    const double ph1 = __shfl_sync(kFull, c.ph_own, 0), ph2 = __shfl_sync(kFull, c.ph_own, 16);
    const cplx p1 = shfl_c(c.wl.prev, kWarpGateLaneO), p2 = shfl_c(c.wl.prev, 16 + kWarpGateLaneO);
    if (lane == 0) {
        st.freq_offset = c.freq_offset; st.pos = c.pos; st.timing_freq = c.timing_freq;
        st.ph1 = ph1; st.ph2 = ph2; st.p1 = p1; st.p2 = p2;
        dstate[stream] = st;
        so.n_sym[stream] = st.n_sym;
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
    // Ask for the largest shared-memory carveout although this kernel needs 2.3 KB per CTA: the L1/shared split of an
    // SM only changes while the SM is empty, and with a small carveout the tracker/decoder CTAs of the previous time
    // tile (22 KB each, second CUDA stream) could not be placed beside the resident demodulator CTAs: they ran after
    // them instead of with them.  The kernel streams its samples through cp.async.cg (L2 only), so it loses nothing.
    int sms = 148, dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (n_streams > 8 * sms) {  // more streams than one wave of the 188-register build
        prefer_max_shared(demod_warp_kernel<16>);
        demod_warp_kernel<16><<<n_streams, 32, 0, st>>>(sb, so, dstate, n_streams, mode, final_flag, afc_alpha, counters);
    } else {
        prefer_max_shared(demod_warp_kernel<8>);
        demod_warp_kernel<8><<<n_streams, 32, 0, st>>>(sb, so, dstate, n_streams, mode, final_flag, afc_alpha, counters);
    }
    return cudaGetLastError();
}

}  // namespace opvd
