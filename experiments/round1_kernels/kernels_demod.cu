// kernels_demod.cu — A1/A3/A4 for sm_100a: int16 I/Q unpack, dual-tone correlator, early-late
// symbol timing recovery, AFC and the batch / streaming call schedule (demod_core.cuh).
//
// A stream is a strictly serial recurrence (symbol n+1's window and LO step depend on symbol n), so
// the parallel axes are streams and, inside one symbol, the two tones and the two halves of the
// 60-sample window.  demod_kernel_t<NT, HALVES> runs L = (2/NT)*HALVES lanes per stream:
//     L=1  both tones on one lane                (fewest instructions per symbol; needs many streams)
//     L=2  one tone per lane
//     L=4  one tone x one window half per lane   (lowest per-symbol latency; small stream counts)
// One warp per CTA, 32/L streams per warp.  Each stream owns a 4-slot ring of 64-sample slots in
// shared memory (+ a mirror of ring slot 0 so the 61-sample window is always contiguous), filled by
// TMA bulk copies (cp.async.bulk + mbarrier complete_tx) issued by the stream's first lane two slots
// ahead of the symbol being demodulated: HBM is read in 256-byte contiguous bursts per stream,
// exactly once, and the LSU never touches global memory on the sample path.  The arithmetic is FP64
// (the reference is FP64 and frame parity is decided at a 3-bit quantiser), restructured to Horner
// polynomials in z = exp(-j*inc) (demod_core.cuh); lanes of one stream exchange their partial gate
// sums, energies, timing error and AFC phase with warp shuffles.  All streams of a warp run the
// symbol loop in lock step (finished streams idle), so the shuffles use the full mask.
#include <cuda_runtime.h>
#include <cstdint>

#include "opvd_kernels.cuh"

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

constexpr int kSlotShift = 6;
constexpr int kSlotSamples = 1 << kSlotShift;               // 64 samples = 256 B per TMA bulk copy
constexpr int kSlotBytes = kSlotSamples * 4;
constexpr int kNumSlots = 4;
constexpr int kRingWords = (kNumSlots + 1) * kSlotSamples;  // + mirror of ring slot 0
constexpr int kRingStrideBytes = kRingWords * 4 + 16;       // 1296 B: odd multiple of 16 B (bank spreading)
constexpr int kMbarBytes = kNumSlots * 8;
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ double shfl_xor_d(double v, int lane_mask) { return __shfl_xor_sync(kFull, v, lane_mask); }

// ------------------------------------------------------------------------------------------------
// One symbol for ONE tone on this lane (NT == 1).  HALVES == 1: the lane evaluates the whole window;
// HALVES == 2: the lane evaluates six of the twelve 5-sample segments and exchanges partial gates.
//
// HALVES == 2 uses a mirrored recurrence so that both halves run the same instruction stream:
//   half 0 (segments 0..5) walks its segments backwards with u = z^5:      R_k = G[5-k] + u R_{k-1}
//   half 1 (segments 6..11) walks forwards with u = conj(z^5):             R_k = G[6+k] + u R_{k-1}
//   gate partials  half 0: (E,O,L) = (R5,R3,R1)        anchored at each gate's first segment
//                  half 1: (E,O,L) = (R1,R3,R5)*z^35   anchored at each gate's last segment
template <int HALVES>
__device__ __forceinline__ Gates tone_gates(const uint32_t* win, cplx z, double f, int half, cplx& z40_out) {
    const TonePowers pw = tone_powers(z);
    z40_out = pw.z40;
    cplx E, O, L;
    if constexpr (HALVES == 1) {
        cplx H[6];
#pragma unroll
        for (int m = 0; m < 6; ++m) {
            double I[10], Q[10];
#pragma unroll
            for (int k = 0; k < 10; ++k) unpack_iq(win[10 * m + k], I[k], Q[k]);
            const cplx a = horner5(I, Q, pw.z), b = horner5(I + 5, Q + 5, pw.z);
            H[m] = cfma(pw.w5, b, a);
        }
        const cplx T01 = cfma(pw.q, H[1], H[0]), T12 = cfma(pw.q, H[2], H[1]), T23 = cfma(pw.q, H[3], H[2]);
        const cplx T34 = cfma(pw.q, H[4], H[3]), T45 = cfma(pw.q, H[5], H[4]);
        E = cfma(pw.q2, T23, T01);
        O = cfma(pw.q2, T34, T12);
        L = cfma(pw.q2, T45, T23);
    } else {
        const int base = half ? 30 : 25, step = half ? 5 : -5;
        const cplx u = half ? cconj(pw.w5) : pw.w5;
        cplx R = {0.0, 0.0}, R1 = {0.0, 0.0}, R3 = {0.0, 0.0};
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const uint32_t* s = win + (base + k * step);
            double I[5], Q[5];
#pragma unroll
            for (int r = 0; r < 5; ++r) unpack_iq(s[r], I[r], Q[r]);
            const cplx G = horner5(I, Q, pw.z);
            R = (k == 0) ? G : cfma(u, R, G);
            if (k == 1) R1 = R;
            if (k == 3) R3 = R;
        }
        // partner scaling: half 1 multiplies by z^35 = z^20 * z^10 * z^5
        const cplx w35 = cmul(cmul(pw.q2, pw.q), pw.w5);
        cplx pe = half ? R1 : R, po = R3, pl = half ? R : R1;
        if (half) { pe = cmul(pe, w35); po = cmul(po, w35); pl = cmul(pl, w35); }
        E = {pe.r + shfl_xor_d(pe.r, 2), pe.i + shfl_xor_d(pe.i, 2)};
        O = {po.r + shfl_xor_d(po.r, 2), po.i + shfl_xor_d(po.i, 2)};
        L = {pl.r + shfl_xor_d(pl.r, 2), pl.i + shfl_xor_d(pl.i, 2)};
    }
    // shifted-window edge terms and the post-sum interpolator
    double sI[6], sQ[6];
    unpack_iq(win[0], sI[0], sQ[0]);
    unpack_iq(win[10], sI[1], sQ[1]);
    unpack_iq(win[20], sI[2], sQ[2]);
    unpack_iq(win[40], sI[3], sQ[3]);
    unpack_iq(win[50], sI[4], sQ[4]);
    unpack_iq(win[60], sI[5], sQ[5]);
    const cplx dE = edge_term(sI[3], sQ[3], sI[0], sQ[0], pw.z40);
    const cplx dO = edge_term(sI[4], sQ[4], sI[1], sQ[1], pw.z40);
    const cplx dL = edge_term(sI[5], sQ[5], sI[2], sQ[2], pw.z40);
    cplx g, h;
    interp_weights(pw.z, f, g, h);
    Gates o;
    o.E = cfma(g, E, cmul(h, dE));
    o.O = cfma(g, O, cmul(h, dO));
    o.L = cfma(g, L, cmul(h, dL));
    return o;
}

// ------------------------------------------------------------------------------------------------
template <int NT, int HALVES>
__global__ void __launch_bounds__(32)
demod_kernel_t(StreamBuffers sb, SoftBuffers so, DemodState* __restrict__ dstate, int n_streams, int mode,
               int final_flag, double afc_alpha, unsigned long long* __restrict__ counters) {
    static_assert(NT == 1 || HALVES == 1, "window halves are only split together with the tones");
    constexpr int L = (2 / NT) * HALVES;  // lanes per stream
    constexpr int SPW = 32 / L;           // streams per warp
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x;
    const int grp = lane / L, sub = lane % L;
    const int tone = (NT == 1) ? (sub & 1) : 0;
    const int half = (HALVES == 2) ? (sub >> 1) : 0;
    const int stream_raw = blockIdx.x * SPW + grp;
    const bool active = stream_raw < n_streams;
    const int stream = active ? stream_raw : n_streams - 1;  // idle groups shadow a real stream, never store

    unsigned char* ring_b = smem + (size_t)grp * kRingStrideBytes;
    const uint32_t* ring = reinterpret_cast<const uint32_t*>(ring_b);
    const uint32_t ring_s = smem_u32(ring_b);
    const uint32_t mbar_s = smem_u32(smem + SPW * kRingStrideBytes + grp * kMbarBytes);
    // 6 doubles per stream: the two tone lanes hand their phase / previous correlation to lane 0 at exit
    double* xch = reinterpret_cast<double*>(smem + SPW * (kRingStrideBytes + kMbarBytes)) + grp * 6;
    const unsigned gmask = (L == 32) ? kFull : (((1u << L) - 1u) << (grp * L));
    (void)xch; (void)gmask;

    if (sub == 0) {
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

    // loop-carried registers.  NT == 2: both tones (DemodRegs).  NT == 1: this lane's tone only.
    DemodRegs r2;
    double freq_offset = st.freq_offset, pos = st.pos, timing_freq = st.timing_freq;
    double ph_own = tone ? st.ph2 : st.ph1, inc_own = 0.0;
    cplx p_own = tone ? st.p2 : st.p1, z_own = {1.0, 0.0};
    if constexpr (NT == 2) regs_from_state(r2, st);
    else lo_step_tone(freq_offset, tone, z_own, inc_own);

    const long long n_sym0 = st.n_sym, origin0 = st.origin;
    bool live = active;
    double call_len_d = (double)st.call_len;   // 0 => the slow path opens the next call
    int origin_rel = (int)(st.origin - row0);
    int a_first = -1, issued = 0, ready = 0;   // slot indices relative to the row (64-sample units)

    while (__any_sync(kFull, live)) {
        double& pos_ref = (NT == 2) ? r2.pos : pos;
        if (live && !((pos_ref + 40.0) + 10.0 < call_len_d)) {  // :221 fails or no call open: slow path
            live = demod_schedule(st, pos_ref, mode, avail, final_flag != 0);
            call_len_d = (double)st.call_len;
            origin_rel = (int)(st.origin - row0);
            if (!live) {
                // the stream is finished for this launch: persist its state now (the lanes keep running
                // the loop in lock step with the rest of the warp, on discarded data)
                if constexpr (NT == 2) {
                    regs_to_state(r2, st);
                } else {
                    if (half == 0) { xch[tone * 3 + 0] = ph_own; xch[tone * 3 + 1] = p_own.r; xch[tone * 3 + 2] = p_own.i; }
                    __syncwarp(gmask);
                    st.freq_offset = freq_offset; st.pos = pos; st.timing_freq = timing_freq;
                    st.ph1 = xch[0]; st.p1 = {xch[1], xch[2]};
                    st.ph2 = xch[3]; st.p2 = {xch[4], xch[5]};
                }
                if (sub == 0) {
                    dstate[stream] = st;
                    unsigned long long dsym = (unsigned long long)(st.n_sym - n_sym0);
                    unsigned long long dsmp = (unsigned long long)(st.origin - origin0);
                    if (st.flags & kFlagDone) dsmp = (unsigned long long)(avail - origin0);
                    if (dsym) atomicAdd(&counters[kCtrSymbols], dsym);
                    if (dsmp) atomicAdd(&counters[kCtrSamples], dsmp);
                }
            }
        }
        const int b = __double2int_rz(pos_ref);          // pos >= 0: truncation == floor (:125)
        const double f = pos_ref - (double)b;
        const int w0 = origin_rel + b - kWinLead;        // row index of window slot 0 (>= -10)
        const int a_lo = w0 >> kSlotShift, a_hi = (w0 + (kWin - 1)) >> kSlotShift;
        __syncwarp();  // every lane is done with the ring slots about to be recycled
        if (live) {
            if (a_first < 0) {  // (re)start: prime the ring
                a_first = a_lo < 0 ? 0 : a_lo;
                issued = ready = a_first;
            }
            // prefetch: slot s overwrites the ring position of slot s-4, which must be behind the window
            while (issued <= a_lo + (kNumSlots - 1) && (issued << kSlotShift) < avail_rel) {
                if (sub == 0) {
                    const int p = issued & (kNumSlots - 1);
                    const int r0 = issued << kSlotShift;
                    const int left = stride_i - r0;
                    const uint32_t bytes = left >= kSlotSamples ? (uint32_t)kSlotBytes : (uint32_t)(left * 4);
                    const uint32_t mb = mbar_s + 8 * p;
                    mbar_expect_tx(mb, p == 0 ? 2 * bytes : bytes);
                    tma_bulk_g2s(ring_s + p * kSlotBytes, row + r0, bytes, mb);
                    if (p == 0) tma_bulk_g2s(ring_s + kNumSlots * kSlotBytes, row + r0, bytes, mb);
                }
                ++issued;
            }
            while (ready <= a_hi) {
                if (ready >= a_first) {
                    const uint32_t mb = mbar_s + 8 * (ready & (kNumSlots - 1));
                    const uint32_t parity = (uint32_t)(((ready - a_first) >> 2) & 1);
                    while (!mbar_try_wait(mb, parity)) {}
                }
                ++ready;
            }
        }
        // idle streams read their (stale) ring at offset 0: finite garbage, results are discarded
        const uint32_t* win = live ? ring + ((a_lo & (kNumSlots - 1)) * kSlotSamples + (w0 - (a_lo << kSlotShift))) : ring;
        const bool first = st.sym_in_call == 0;
        double soft;
        if constexpr (NT == 2) {
            soft = demod_symbol(r2, win, f, first, afc_alpha);
        } else {
            cplx z40;
            Gates g = tone_gates<HALVES>(win, z_own, f, half, z40);
            if (first) {  // early-gate clamp at the start of a call (:237)
                const cplx fix = first_symbol_fix(win, f, z_own);
                g.E.r -= fix.r; g.E.i -= fix.i;
            }
            const double e_own = cnorm(g.O);
            const double e_oth = shfl_xor_d(e_own, 1);
            const double e1 = tone ? e_oth : e_own, e2 = tone ? e_own : e_oth;
            soft = e2 - e1;                                   // :268
            const bool dom = (tone == 0) == (e1 > e2);        // this lane holds the dominant tone (:272)
            const double ted_own = ted_from_gates(g.E, g.L);
            const double ted_oth = shfl_xor_d(ted_own, 1);
            const double timing_adj = timing_loop(timing_freq, dom ? ted_own : ted_oth);
            const cplx n_own = cmul(g.O, cconj(z40));
            const double pd_own = first ? 0.0 : afc_phase(g.O, p_own, ph_own);
            const double pd_oth = shfl_xor_d(pd_own, 1);
            if (!first) afc_loop(freq_offset, dom ? pd_own : pd_oth, afc_alpha);  // :289-307
            ph_own = wrap_phase(fma(40.0, inc_own, ph_own));
            if (!first) lo_step_tone(freq_offset, tone, z_own, inc_own);
            p_own = n_own;
            pos += 40.0 + timing_adj;                          // :313
        }
        if (live) {
            if (sub == 0) soft_row[st.n_sym] = soft;
            st.n_sym++;
            st.sym_in_call++;
        }
    }

    // every stream persisted its state when it left the loop (see the slow path above)
}

template <int NT, int HALVES>
static cudaError_t launch_t(const StreamBuffers& sb, const SoftBuffers& so, DemodState* dstate, int n_streams,
                            int mode, int final_flag, double afc_alpha, unsigned long long* counters,
                            cudaStream_t st) {
    constexpr int L = (2 / NT) * HALVES, SPW = 32 / L;
    const size_t smem = (size_t)SPW * (kRingStrideBytes + kMbarBytes + 6 * sizeof(double));
    cudaError_t e = cudaFuncSetAttribute(demod_kernel_t<NT, HALVES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int grid = (n_streams + SPW - 1) / SPW;
    demod_kernel_t<NT, HALVES><<<grid, 32, smem, st>>>(sb, so, dstate, n_streams, mode, final_flag, afc_alpha, counters);
    return cudaGetLastError();
}

int demod_auto_lanes(int n_streams) {
    // Small banks are bound by the per-symbol latency of the serial recurrence: one warp per stream (12 resident
    // warps per SM at the kernel's register count; it peaks at 82 Gsample/s with 1,776 streams).  Larger banks: the
    // batched kernel (32 streams per CTA), which spends ~5x fewer instructions per stream and symbol and scales with
    // the stream count (108 Gsample/s at 4,096 streams, 320 at 18,944).  Measured crossover (profiles/README.md,
    // round-1 "f" sweep): ~16 streams per SM.  The lane kernels (1, 2, 4) and the pipelined kernel (128) stay
    // selectable for comparison.
    int sms = 148, dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if ((long long)n_streams <= 16ll * sms) return 32;
    return 64;
}

cudaError_t launch_demod(const StreamBuffers& sb, const SoftBuffers& so, DemodState* dstate, int n_streams,
                         int mode, int final_flag, double afc_alpha, int lanes_per_stream,
                         unsigned long long* counters, cudaStream_t st) {
    if (n_streams <= 0) return cudaSuccess;
    int L = lanes_per_stream > 0 ? lanes_per_stream : demod_auto_lanes(n_streams);
    if (L >= 128) return launch_demod_pipe(sb, so, dstate, n_streams, mode, final_flag, afc_alpha, counters, st);
    if (L >= 96) return launch_demod_bank(sb, so, dstate, n_streams, mode, final_flag, afc_alpha, counters, st);
    if (L >= 64) return launch_demod_batch(sb, so, dstate, n_streams, mode, final_flag, afc_alpha, counters, st);
    if (L >= 32) return launch_demod_warp(sb, so, dstate, n_streams, mode, final_flag, afc_alpha, counters, st);
    if (L >= 4) return launch_t<1, 2>(sb, so, dstate, n_streams, mode, final_flag, afc_alpha, counters, st);
    if (L >= 2) return launch_t<1, 1>(sb, so, dstate, n_streams, mode, final_flag, afc_alpha, counters, st);
    return launch_t<2, 1>(sb, so, dstate, n_streams, mode, final_flag, afc_alpha, counters, st);
}

}  // namespace opvd
