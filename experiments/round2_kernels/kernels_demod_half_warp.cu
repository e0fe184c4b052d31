// kernels_demod_half.cu — A1/A3/A4 for sm_100a, HALF-WARP-PER-STREAM variant (demod_warp_core.cuh).
//
// The warp-per-stream kernel (kernels_demod_warp.cu) is bound by the per-symbol latency of one warp, and that latency
// grows by a fifth as soon as two of its warps share an SM sub-partition (1,019 -> 1,240 cycles per symbol: they queue
// for the FP64 pipe's operand fetch, profiles/ncu_warp_r02_u_roles.txt).  A B200 has 592 sub-partitions; the headline
// bank has 1,024 streams.  This kernel keeps ONE warp per sub-partition for banks of 593..1,184 streams by giving each
// stream a half-warp: lane p = lane & 15 owns the same five window slots as in the warp kernel, but for BOTH tones
// (two WarpLane records per lane), the gate sums are formed by 16-wide shuffles inside the half, and the two halves of
// a warp run two unrelated streams through one instruction stream.  The second tone's Horner chain, shuffles and gate
// arithmetic are independent of the first tone's and issue in its dependency bubbles; one instruction now serves two
// streams, so the FP64 pipe sees half as many warp instructions per stream.
//
// Both halves execute the same flat loop, one symbol per iteration; everything that happens at different times in the
// two streams (call boundaries :1012-1113, ring refills, the end of a stream) is a short divergent section inside the
// iteration.  (Nested loops, as in the warp kernel, would park the half that leaves the inner loop at the loop's
// reconvergence point until the other half leaves it too, one chunk later.)  The per-lane arithmetic is exactly the
// warp kernel's (same functions, same operation order), so the results are bit-identical to it.
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
constexpr int kSlotSamples = 1 << kSlotShift;  // 64 samples = 256 B per slot: 16 lanes x 16 bytes
constexpr int kSlotBytes = kSlotSamples * 4;
constexpr int kNumSlots = 8;
constexpr int kRingSamples = kNumSlots * kSlotSamples;     // 512 samples = 2 KB per stream
constexpr int kRingWords = kRingSamples + kSlotSamples;    // + mirror of ring slot 0: windows never wrap
constexpr int kRingMask = kRingSamples - 1;
constexpr unsigned kFull = 0xffffffffu;

// Everything one half-warp carries through the symbol loop of its stream (every lane holds a copy of the uniform part).
struct HalfCtx {
    FastMathTable K;     // polynomial / loop constants pinned in registers
    WarpLane w1, w2;     // this lane's slots for tone F1 and tone F2
    cplx RP1, RP2;       // R * prev per tone (gate lane O): O_n * conj(prev) = X * conj(RP)
    bool prev_zero1, prev_zero2;
    double freq_offset, pos, timing_freq, ph1, ph2, afc_alpha;
    unsigned hmask;      // the 16 lanes of this half
    const uint32_t* ring;
    uint32_t ring_s;
    const uint32_t* row;
    int row_base_off, row_wrap, rel_end, issued_off;
    int avail_rel, origin_rel;
    int issued_s;        // samples [.., issued_s) of the row have been requested (multiple of 64); < 0: ring not primed
    int hl, lane_slot;
    double* soft_row;
    int soft_idx, soft_wrap, n_new;

    __device__ __forceinline__ cplx shfl_down_c(cplx v, int d) const {
        return {__shfl_down_sync(hmask, v.r, d, 16), __shfl_down_sync(hmask, v.i, d, 16)};
    }
    __device__ __forceinline__ cplx shfl_c(cplx v, int src) const {
        return {__shfl_sync(hmask, v.r, src, 16), __shfl_sync(hmask, v.i, src, 16)};
    }
    __device__ __forceinline__ double shfl_d(double v, int src) const { return __shfl_sync(hmask, v, src, 16); }

    __device__ __forceinline__ void set_view(const RowView& v) {
        row = v.row; row_base_off = v.base_off; row_wrap = v.wrap; rel_end = v.rel_end; issued_off = 0;
    }
    // request the next 64-sample slot: the 16 lanes of the half copy 16 bytes each
    __device__ __forceinline__ void issue_slot() {
        const int p = (issued_s >> kSlotShift) & (kNumSlots - 1);
        const int off = issued_s + 4 * hl;
        if (off + 4 <= rel_end) {
            const uint32_t* src = row + issued_off + 4 * hl;
            cp_async16(ring_s + p * kSlotBytes + 16 * hl, src);
            if (p == 0) cp_async16(ring_s + kNumSlots * kSlotBytes + 16 * hl, src);
        }
        issued_s += kSlotSamples;
        issued_off += kSlotSamples;
        if (issued_off >= row_wrap) issued_off -= row_wrap;
    }
    // once per symbol: keep seven slots requested ahead of the window starting at row index w0 (kernels_demod_warp.cu)
    __device__ __forceinline__ void ring_maintain(int w0) {
        if (w0 + (kNumSlots - 1) * kSlotSamples >= issued_s && issued_s < avail_rel) {
            __syncwarp(hmask);  // every lane of the half is done with the slot about to be recycled
            issue_slot();
            ph1 = warp_wrap_phase(ph1, K);  // only read in the signed-zero corner: wrapped here, every ~1.6 symbols
            ph2 = warp_wrap_phase(ph2, K);
        }
        cp_async_commit();
        cp_async_wait<4>();
        __syncwarp(hmask);
    }
    __device__ __forceinline__ void ring_prime(int w0) {
        issued_s = w0 < 0 ? 0 : (w0 & ~(kSlotSamples - 1));
        issued_off = row_base_off + issued_s;
        while (issued_off >= row_wrap) issued_off -= row_wrap;
        while (w0 + (kNumSlots - 1) * kSlotSamples >= issued_s && issued_s < avail_rel) issue_slot();
        cp_async_commit();
        cp_async_wait<0>();
        __syncwarp(hmask);
    }
    __device__ __forceinline__ void load_window(int b, double (&I5)[5], double (&Q5)[5]) const {
        const uint32_t* src = ring + (((origin_rel + b - kWinLead) & kRingMask) + lane_slot);
        uint32_t s5[5];
#pragma unroll
        for (int r = 0; r < 5; ++r) s5[r] = src[r];
#pragma unroll
        for (int r = 0; r < 5; ++r) unpack_iq(s5[r], I5[r], Q5[r]);
    }

    // interpolated gate of one tone on the gate-owner lanes (p = 0, 2, 4)
    __device__ __forceinline__ cplx gates(const WarpLane& w, double f, const LanePartial& lp) const {
        const cplx Fh = shfl_down_c(lp.F, 8);
        cplx acc = lp.W;
#pragma unroll
        for (int d = 1; d <= 4; d <<= 1) {
            const cplx o = shfl_down_c(acc, d);
            acc = {acc.r + o.r, acc.i + o.i};
        }
        return warp_lane_gate(w, f, acc, Fh, lp.F);
    }

    // One symbol at integer position b = trunc(pos), whose five lane samples are in I5/Q5.  Updates pos, loads and
    // converts the NEXT symbol's samples and returns its b, then runs the AFC chain.  first: first symbol of a
    // demodulate() call (early-gate clamp :237, no AFC update :289).
    __device__ __forceinline__ int symbol(int b, bool first, double (&I5)[5], double (&Q5)[5]) {
        const double f = pos - (double)b;
        const LanePartial lp1 = warp_lane_partial_d(w1, I5, Q5), lp2 = warp_lane_partial_d(w2, I5, Q5);
        cplx X1 = gates(w1, f, lp1), X2 = gates(w2, f, lp2);
        if (first) {
            const uint32_t* win = ring + ((origin_rel + b - kWinLead) & kRingMask);
            const cplx fix1 = first_symbol_fix_w([&](int k) { return win[k]; }, f, w1.z);
            const cplx fix2 = first_symbol_fix_w([&](int k) { return win[k]; }, f, w2.z);
            if (w1.p == kWarpGateLaneE) { X1.r -= fix1.r; X1.i -= fix1.i; X2.r -= fix2.r; X2.i -= fix2.i; }
        }
        const double nrm1 = cnorm(X1), nrm2 = cnorm(X2);
        const double e1 = shfl_d(nrm1, kWarpGateLaneO), e2 = shfl_d(nrm2, kWarpGateLaneO);
        const double eE1 = shfl_d(nrm1, kWarpGateLaneE), eL1 = shfl_d(nrm1, kWarpGateLaneL);
        const double eE2 = shfl_d(nrm2, kWarpGateLaneE), eL2 = shfl_d(nrm2, kWarpGateLaneL);
        bool tone1;
        const double soft = warp_uniform_timing(e1, e2, eE1, eL1, eE2, eL2, timing_freq, pos, tone1, K);
        if (hl == 0) soft_row[soft_idx] = soft;
        if (++soft_idx == soft_wrap) soft_idx = 0;
        ++n_new;

        const int b_next = __double2int_rz(pos);  // pos >= 0: truncation == floor (:125)
        load_window(b_next, I5, Q5);

        // ---- AFC (:289-307): the phase detector of the dominant tone, on gate lane O
        const bool x_zero1 = nrm1 == 0.0, x_zero2 = nrm2 == 0.0;
        double pd_own = 0.0;
        if (!first) {
            const cplx X = tone1 ? X1 : X2, RP = tone1 ? RP1 : RP2;
            const bool corner = tone1 ? (x_zero1 || prev_zero1) : (x_zero2 || prev_zero2);
            const double xr = fma(X.r, RP.r, X.i * RP.i);
            const double xi = fma(X.i, RP.r, -(X.r * RP.i));
            pd_own = atan2_fast(xi, xr, K);
            if (corner) {
                const cplx R = tone1 ? w1.R : w2.R, prev = tone1 ? w1.prev : w2.prev;
                pd_own = afc_phase_corner(cmul(X, cconj(R)), prev, tone1 ? ph1 : ph2);
            }
        }
        const cplx z50a = shfl_c(w1.R, 10), z50b = shfl_c(w2.R, 10);  // R of lane p = 10 is z^50 = z^10 * z^40
        w1.prev = cmul(X1, cconj(z50a));                               // :309-310, in the next symbol's phase frame
        w2.prev = cmul(X2, cconj(z50b));
        prev_zero1 = x_zero1; prev_zero2 = x_zero2;
        ph1 = fma(40.0, w1.inc, ph1);                                  // :250-262 (wrapped in ring_maintain)
        ph2 = fma(40.0, w2.inc, ph2);
        if (!first) {
            const double pd = shfl_d(pd_own, kWarpGateLaneO);
            warp_afc_loop(freq_offset, pd, afc_alpha, K);
            warp_lane_lo_fast(w1, freq_offset, K);
            warp_lane_lo_fast(w2, freq_offset, K);
        }
        RP1 = cmul(w1.R, w1.prev);
        RP2 = cmul(w2.R, w2.prev);
        return b_next;
    }
};

}  // namespace

__global__ void __launch_bounds__(32)
demod_half_kernel(StreamBuffers sb, SoftBuffers so, DemodState* __restrict__ dstate, int n_streams, int mode,
                  int final_flag, double afc_alpha, unsigned long long* __restrict__ counters) {
    __shared__ __align__(128) uint32_t ring_sm[2][kRingWords];
    const int lane = threadIdx.x, half = lane >> 4;
    const int stream_raw = 2 * blockIdx.x + half;
    const bool valid = stream_raw < n_streams;
    const int stream = valid ? stream_raw : n_streams - 1;

    HalfCtx c;
    c.K = load_table_pinned();
    c.hl = lane & 15;
    c.hmask = 0xffffu << (16 * half);
    c.ring = ring_sm[half];
    c.ring_s = smem_u32(ring_sm[half]);

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

    warp_lane_init(c.w1, c.hl, c.K);        // tone F1, p = hl
    warp_lane_init(c.w2, 16 + c.hl, c.K);   // tone F2, p = hl
    warp_lane_lo(c.w1, st.freq_offset);     // general version: a -o offset may exceed the fast range
    warp_lane_lo(c.w2, st.freq_offset);
    c.w1.prev = st.p1; c.w2.prev = st.p2;
    c.prev_zero1 = st.p1.r == 0.0 && st.p1.i == 0.0;
    c.prev_zero2 = st.p2.r == 0.0 && st.p2.i == 0.0;
    c.RP1 = cmul(c.w1.R, c.w1.prev);
    c.RP2 = cmul(c.w2.R, c.w2.prev);
    c.lane_slot = 5 * (c.w1.p > 12 ? 12 : c.w1.p);
    c.freq_offset = st.freq_offset;
    c.pos = st.pos;
    c.timing_freq = st.timing_freq;
    c.ph1 = st.ph1; c.ph2 = st.ph2;
    c.issued_s = -1;

    const long long n_sym0 = st.n_sym, origin0 = st.origin;
    bool live = valid;
    bool in_call = false;   // a demodulate() call is open and its next symbol fits (:221)
    int b = 0, call_len_i = 0;
    double call_len_d = 0.0;
    double I5[5], Q5[5];
#pragma unroll
    for (int r = 0; r < 5; ++r) { I5[r] = 0.0; Q5[r] = 0.0; }

    // flat loop: one symbol of each half per iteration
    while (__any_sync(kFull, live)) {
        if (live) {
            bool first = false;
            if (!in_call) {
                // ---- call boundary (uniform within the half): close / open demodulate() calls (:1012-1113)
                st.n_sym = n_sym0 + c.n_new;
                if (!demod_schedule(st, c.pos, mode, avail, final_flag != 0)) {
                    live = false;
                } else {
                    c.origin_rel = (int)(st.origin - row0);
                    call_len_i = (int)st.call_len;
                    call_len_d = (double)st.call_len;
                    b = __double2int_rz(c.pos);  // pos >= 0: truncation == floor (:125)
                    if (c.issued_s < 0) c.ring_prime(c.origin_rel + b - kWinLead);  // first symbol of this launch
                    else c.ring_maintain(c.origin_rel + b - kWinLead);
                    c.load_window(b, I5, Q5);
                    first = st.sym_in_call == 0;
                    in_call = true;
                }
            } else {
                c.ring_maintain(c.origin_rel + b - kWinLead);
            }
            if (live) {
                b = c.symbol(b, first, I5, Q5);
                st.sym_in_call = 2;  // any non-zero value: the open call has produced symbols
                // while (pos + 50 < N) (:221); the integer test is a conservative shortcut
                if (b + 52 >= call_len_i && !((c.pos + 40.0) + 10.0 < call_len_d)) in_call = false;
            }
        }
    }

    // persist the stream's state
    c.ph1 = warp_wrap_phase(c.ph1, c.K);
    c.ph2 = warp_wrap_phase(c.ph2, c.K);
    const cplx p1 = c.shfl_c(c.w1.prev, kWarpGateLaneO), p2 = c.shfl_c(c.w2.prev, kWarpGateLaneO);
    if (c.hl == 0 && valid) {
        st.freq_offset = c.freq_offset; st.pos = c.pos; st.timing_freq = c.timing_freq;
        st.ph1 = c.ph1; st.ph2 = c.ph2; st.p1 = p1; st.p2 = p2;
        dstate[stream] = st;
        so.n_sym[stream] = st.n_sym;
        unsigned long long dsym = (unsigned long long)(st.n_sym - n_sym0);
        unsigned long long dsmp = (unsigned long long)(st.origin - origin0);
        if (st.flags & kFlagDone) dsmp = (unsigned long long)(avail - origin0);
        if (dsym) atomicAdd(&counters[kCtrSymbols], dsym);
        if (dsmp) atomicAdd(&counters[kCtrSamples], dsmp);
    }
}

cudaError_t launch_demod_half(const StreamBuffers& sb, const SoftBuffers& so, DemodState* dstate, int n_streams,
                              int mode, int final_flag, double afc_alpha, unsigned long long* counters,
                              cudaStream_t st) {
    // largest shared-memory carveout, for the reason given in launch_demod_warp (tracker/decoder CTAs beside this kernel)
    cudaFuncSetAttribute(demod_half_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    demod_half_kernel<<<(n_streams + 1) / 2, 32, 0, st>>>(sb, so, dstate, n_streams, mode, final_flag, afc_alpha, counters);
    return cudaGetLastError();
}

}  // namespace opvd
