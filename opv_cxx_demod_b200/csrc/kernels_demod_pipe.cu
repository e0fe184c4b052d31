// kernels_demod_pipe.cu — A1/A3/A4 for sm_100a, PIPELINED batched variant for large channel banks.
// Same arithmetic as kernels_demod_batch.cu (demod_batch_core.cuh); different schedule.
//
// The batched kernel alternates a window phase and a loop phase per symbol, so at any time about half
// of its warps wait at a barrier, and its window warps convert every sample once per tone.  Here a CTA
// of 128 threads owns TWO groups of 32 streams and its four warps are specialised:
//   warps W0, W1  window workers: half h of the 60-sample window for BOTH tones (samples are loaded and
//                 converted once), then — after a two-warp named barrier — W_h combines the halves of
//                 tone h and publishes its gate energies / on-time sum;
//   warp  T       timing chain of the group whose window finished one period earlier (soft decision,
//                 TED, timing loop, call schedule, soft store) and that group's ring staging (cp.async
//                 straight from HBM into the transposed ring, no registers, a full symbol to land);
//   warp  A       AFC chain of the same group (phase detector, AFC loop, LO steps of its next symbol).
// Period p: window(group p & 1) runs while loop(group ~p & 1) runs; one CTA barrier per period.  Every
// warp is busy in every period, and the instruction count per stream and symbol drops by ~30 %.
// Lane = stream everywhere, so there is still no intra-warp exchange.
#include <cuda_runtime.h>
#include <cstdint>

#include "demod_batch_core.cuh"
#include "demod_warp_core.cuh"  // first_symbol_fix_w
#include "opvd_kernels.cuh"

namespace opvd {

namespace {

constexpr int kSpc = 32;            // streams per group
constexpr int kThreads = 128;
constexpr int kRingRows = 256;      // samples per stream resident in shared memory (power of two)
constexpr int kMirrorRows = 64;     // rows 0..63 repeated after row 255: a 61-row window never wraps
constexpr int kRows = kRingRows + kMirrorRows;
constexpr int kSub = 8;             // samples per 32-byte sector
constexpr int kStageAll = 6 * kSub; // samples staged per stream and symbol

struct __align__(16) GroupSmem {
    uint32_t ring[kRows][kSpc];     // transposed sample ring, 40 KB
    double2 part[2][2][3][kSpc];    // [tone][half][E,O,L] interpolated partial gates
    double tg[2][7][kSpc];          // per tone: eE, eO, eL, O.r, O.i, z40.r, z40.i
    double zq[2][4][kSpc];          // [tone][z.r, z.i, q.r, q.i]: LO steps of the group's next window
    double frac[kSpc];              // interpolation fraction of the next window
    int w0[kSpc];                   // row-relative sample index of slot 0 of the next window
    int live[kSpc], first[kSpc];    // next window: stream has a symbol / it is the first of a call
    int sym_live[kSpc], sym_first[kSpc];  // the same two for the symbol the loop warps are working on
    int any_live;                   // some stream of the group has a next window
    int ran;                        // the window workers processed the group in the previous period
};
struct __align__(16) PipeSmem {
    GroupSmem g[2];
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void pair_barrier() { asm volatile("bar.sync 1, 64;" ::: "memory"); }  // W0 + W1

__device__ __noinline__ cplx first_fix_cold(const uint32_t* win, double f, cplx z) {
    return first_symbol_fix_w([&](int kk) { return win[kk * kSpc]; }, f, z);
}
__device__ __noinline__ bool schedule_cold(DemodState& st, int mode, long long avail, bool final_flag) {
    double pos = st.pos;
    const bool live = demod_schedule(st, pos, mode, avail, final_flag);
    st.pos = pos;
    return live;
}

// ---- ring staging: up to 48 samples of one stream, asynchronously (timing warp)
__device__ __forceinline__ void stage_async(GroupSmem& sm, int s, const uint32_t* __restrict__ row, int stride, int& fill,
                                            int w0) {
    if (fill + kStageAll <= w0 + kRingRows) {
#pragma unroll
        for (int c = 0; c < kStageAll / kSub; ++c) {
            const int idx = fill + kSub * c;
            if (idx + kSub <= stride) {
                const int r = idx & (kRingRows - 1);
                const uint32_t dst = smem_u32(&sm.ring[r][s]);
#pragma unroll
                for (int j = 0; j < kSub; ++j) cp_async4(dst + j * (kSpc * 4), row + idx + j);
                if (r < kMirrorRows) {
#pragma unroll
                    for (int j = 0; j < kSub; ++j) cp_async4(dst + (kRingRows + j) * (kSpc * 4), row + idx + j);
                }
            } else if (idx < stride) {  // last, partial sector of a row
                for (int j = 0; idx + j < stride; ++j) {
                    const int r = (idx + j) & (kRingRows - 1);
                    cp_async4(smem_u32(&sm.ring[r][s]), row + idx + j);
                    if (r < kMirrorRows) cp_async4(smem_u32(&sm.ring[kRingRows + r][s]), row + idx + j);
                }
            }
        }
        fill += kStageAll;
    }
}

// ---- window worker: half `half` of the window, both tones; then the halves of tone `half` are combined
__device__ __forceinline__ void window_role(GroupSmem& sm, int s, int half) {
    if (!sm.any_live) {  // uniform
        if (half == 0 && s == 0) sm.ran = 0;
        return;
    }
    const int lv = sm.live[s], w0 = sm.w0[s];
    const bool first = sm.first[s] != 0;
    const double f = sm.frac[s];
    if (half == 0) {
        sm.sym_live[s] = lv;
        sm.sym_first[s] = first;
        if (s == 0) sm.ran = 1;
    }
    const uint32_t* win = &sm.ring[w0 & (kRingRows - 1)][s];
    if (lv) {
        const uint32_t* src = win + 30 * half * kSpc;
        double I[31], Q[31];  // slots 30h .. 30h+29, and slot 60 for the late gate's edge term (h = 1)
#pragma unroll
        for (int j = 0; j < 30; ++j) unpack_iq_mixed(src[j * kSpc], I[j], Q[j]);
        I[30] = 0.0; Q[30] = 0.0;
        if (half) unpack_iq(src[30 * kSpc], I[30], Q[30]);
#pragma unroll
        for (int tone = 0; tone < 2; ++tone) {
            const cplx z = {sm.zq[tone][0][s], sm.zq[tone][1][s]}, q = {sm.zq[tone][2][s], sm.zq[tone][3][s]};
            const HalfGates g = batch_half_gates(I, Q, z, q, f, half);
            sm.part[tone][half][0][s] = make_double2(g.E.r, g.E.i);
            sm.part[tone][half][1][s] = make_double2(g.O.r, g.O.i);
            sm.part[tone][half][2][s] = make_double2(g.L.r, g.L.i);
        }
    }
    pair_barrier();
    if (lv) {
        const int tone = half;
        HalfGates a, b;
        double2 v;
        v = sm.part[tone][0][0][s]; a.E = {v.x, v.y}; v = sm.part[tone][0][1][s]; a.O = {v.x, v.y};
        v = sm.part[tone][0][2][s]; a.L = {v.x, v.y};
        v = sm.part[tone][1][0][s]; b.E = {v.x, v.y}; v = sm.part[tone][1][1][s]; b.O = {v.x, v.y};
        v = sm.part[tone][1][2][s]; b.L = {v.x, v.y};
        ToneLo t;
        t.z = {sm.zq[tone][0][s], sm.zq[tone][1][s]};
        t.q = {sm.zq[tone][2][s], sm.zq[tone][3][s]};
        t.inc = 0.0;
        cplx fix = {0.0, 0.0};
        if (first) fix = first_fix_cold(win, f, t.z);  // early-gate clamp (:237), once per call
        const ToneGates g = batch_finish_tone(a, b, t, fix);
        sm.tg[tone][0][s] = g.eE; sm.tg[tone][1][s] = g.eO; sm.tg[tone][2][s] = g.eL;
        sm.tg[tone][3][s] = g.O.r; sm.tg[tone][4][s] = g.O.i; sm.tg[tone][5][s] = g.z40.r; sm.tg[tone][6][s] = g.z40.i;
    }
}

// ---- timing warp state of one group
struct TimingState {
    DemodState st;  // local memory: only the out-of-line scheduler touches it
    double pos, timing_freq, call_len_d;
    long long avail, n_sym0, origin0;
    double* soft_row;
    double* soft_ptr;
    const uint32_t* row;
    int origin_rel, sym_in_call, w0, fill, stride;
    bool live, valid;
};

__device__ __forceinline__ void timing_publish(GroupSmem& sm, int s, TimingState& t) {
    if (t.live) {
        const int b = __double2int_rz(t.pos);  // pos >= 0: truncation == floor (:125)
        t.w0 = t.origin_rel + b - kWinLead;
        sm.w0[s] = t.w0;
        sm.frac[s] = t.pos - (double)b;
        sm.first[s] = t.sym_in_call == 0;
    }
    sm.live[s] = t.live ? 1 : 0;
    const int any = __any_sync(0xffffffffu, t.live);
    if (s == 0) sm.any_live = any;
}

__device__ __forceinline__ void timing_init(GroupSmem& sm, int s, TimingState& t, const StreamBuffers& sb,
                                            const SoftBuffers& so, const DemodState* dstate, int stream, bool valid,
                                            int mode, int final_flag) {
    t.valid = valid;
    t.st = dstate[stream];
    t.avail = sb.avail[stream];
    t.row = sb.iq + (long long)stream * sb.stride;
    t.stride = (int)sb.stride;
    t.soft_row = so.soft + (long long)stream * so.stride - so.base;
    t.soft_ptr = t.soft_row + t.st.n_sym;
    t.n_sym0 = t.st.n_sym; t.origin0 = t.st.origin;
    t.timing_freq = t.st.timing_freq;
    t.live = valid && schedule_cold(t.st, mode, t.avail, final_flag != 0);
    t.pos = t.st.pos;
    t.sym_in_call = t.st.sym_in_call;
    t.call_len_d = (double)t.st.call_len;
    t.origin_rel = (int)(t.st.origin - sb.row_base);
    t.w0 = 0;
    timing_publish(sm, s, t);
    t.fill = (t.w0 < 0 ? 0 : t.w0) & ~(kSub - 1);
    if (s == 0) sm.ran = 0;
}

__device__ __forceinline__ void timing_role(GroupSmem& sm, int s, TimingState& t, long long row0, int mode, int final_flag) {
    if (!sm.ran) {  // uniform: nothing to do for this group, but keep the cp.async groups alternating between the
        cp_async_wait<1>();  // two stream groups (the other group's wait<1> counts on it)
        cp_async_commit();
        return;
    }
    if (sm.sym_live[s]) {
        const double soft = batch_timing(sm.tg[0][1][s], sm.tg[1][1][s], sm.tg[0][0][s], sm.tg[0][2][s], sm.tg[1][0][s],
                                         sm.tg[1][2][s], t.timing_freq, t.pos, g_fm);
        *t.soft_ptr++ = soft;
        t.sym_in_call = 1;  // any non-zero value: the open call has produced symbols
        if (!((t.pos + 40.0) + 10.0 < t.call_len_d)) {  // :221 fails: close the call, maybe open the next
            t.st.n_sym = (long long)(t.soft_ptr - t.soft_row);
            t.st.sym_in_call = t.sym_in_call;
            t.st.pos = t.pos;
            t.live = schedule_cold(t.st, mode, t.avail, final_flag != 0);
            t.pos = t.st.pos;
            t.sym_in_call = t.st.sym_in_call;
            t.call_len_d = (double)t.st.call_len;
            t.origin_rel = (int)(t.st.origin - row0);
        }
    }
    const int w0_sym = t.w0;  // window of the symbol just finished: everything older is dead
    timing_publish(sm, s, t);
    // staging: the batch requested two visits ago has landed; request the next one
    cp_async_wait<1>();
    if (sm.sym_live[s]) stage_async(sm, s, t.row, t.stride, t.fill, w0_sym);
    cp_async_commit();
}

__device__ __forceinline__ void timing_finish(TimingState& t, DemodState* dstate, int stream,
                                              unsigned long long* counters) {
    if (!t.valid) return;
    t.st.n_sym = (long long)(t.soft_ptr - t.soft_row);
    t.st.sym_in_call = t.sym_in_call;
    t.st.pos = t.pos; t.st.timing_freq = t.timing_freq;
    dstate[stream] = t.st;
    unsigned long long dsym = (unsigned long long)(t.st.n_sym - t.n_sym0);
    unsigned long long dsmp = (unsigned long long)(t.st.origin - t.origin0);
    if (t.st.flags & kFlagDone) dsmp = (unsigned long long)(t.avail - t.origin0);
    if (dsym) atomicAdd(&counters[kCtrSymbols], dsym);
    if (dsmp) atomicAdd(&counters[kCtrSamples], dsmp);
}

// ---- AFC warp state of one group
struct AfcState {
    BatchAfc afc;
    double inc1, inc2;  // LO phase steps of the symbol in flight
};

__device__ __forceinline__ void afc_publish(GroupSmem& sm, int s, const ToneLo& t1, const ToneLo& t2) {
    sm.zq[0][0][s] = t1.z.r; sm.zq[0][1][s] = t1.z.i; sm.zq[0][2][s] = t1.q.r; sm.zq[0][3][s] = t1.q.i;
    sm.zq[1][0][s] = t2.z.r; sm.zq[1][1][s] = t2.z.i; sm.zq[1][2][s] = t2.q.r; sm.zq[1][3][s] = t2.q.i;
}
__device__ __forceinline__ void afc_init(GroupSmem& sm, int s, AfcState& a, const DemodState* dstate, int stream) {
    const DemodState* d = dstate + stream;
    a.afc.freq_offset = d->freq_offset; a.afc.ph1 = d->ph1; a.afc.ph2 = d->ph2; a.afc.p1 = d->p1; a.afc.p2 = d->p2;
    ToneLo t1, t2;
    batch_lo(a.afc.freq_offset, t1, t2);  // general version: a -o offset may exceed the fast range
    a.inc1 = t1.inc; a.inc2 = t2.inc;
    afc_publish(sm, s, t1, t2);
}
__device__ __forceinline__ void afc_role(GroupSmem& sm, int s, AfcState& a, double afc_alpha) {
    if (!sm.ran) return;  // uniform
    if (sm.sym_live[s]) {
        const bool first = sm.sym_first[s] != 0;
        const cplx O1 = {sm.tg[0][3][s], sm.tg[0][4][s]}, z40_1 = {sm.tg[0][5][s], sm.tg[0][6][s]};
        const cplx O2 = {sm.tg[1][3][s], sm.tg[1][4][s]}, z40_2 = {sm.tg[1][5][s], sm.tg[1][6][s]};
        batch_afc(a.afc, O1, z40_1, sm.tg[0][1][s], O2, z40_2, sm.tg[1][1][s], a.inc1, a.inc2, first, afc_alpha, g_fm);
        if (!first) {
            ToneLo t1, t2;
            batch_lo_fast(a.afc.freq_offset, t1, t2, g_fm);  // |freq_offset| <= 2 kHz after the AFC clamp
            a.inc1 = t1.inc; a.inc2 = t2.inc;
            afc_publish(sm, s, t1, t2);
        }
    }
}
__device__ __forceinline__ void afc_finish(const AfcState& a, DemodState* dstate, int stream) {
    DemodState* d = dstate + stream;
    d->freq_offset = a.afc.freq_offset; d->ph1 = a.afc.ph1; d->ph2 = a.afc.ph2; d->p1 = a.afc.p1; d->p2 = a.afc.p2;
}

}  // namespace

__global__ void __launch_bounds__(kThreads, 2)
demod_pipe_kernel(StreamBuffers sb, SoftBuffers so, DemodState* __restrict__ dstate, int n_streams, int mode,
                  int final_flag, double afc_alpha, unsigned long long* __restrict__ counters) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PipeSmem& sm = *reinterpret_cast<PipeSmem*>(smem_raw);
    // roles rotate with the CTA index so that co-resident CTAs spread the heavy roles over the SM sub-partitions
    const int s = threadIdx.x & 31, role = ((threadIdx.x >> 5) + blockIdx.x) & 3;  // 0,1: window halves; 2: timing; 3: AFC
    const int raw0 = blockIdx.x * (2 * kSpc) + s, raw1 = raw0 + kSpc;
    const bool valid0 = raw0 < n_streams, valid1 = raw1 < n_streams;
    const int stream0 = valid0 ? raw0 : n_streams - 1, stream1 = valid1 ? raw1 : n_streams - 1;
    const long long row0 = sb.row_base;

    TimingState t0, t1;
    AfcState a0, a1;
    if (role == 2) {
        timing_init(sm.g[0], s, t0, sb, so, dstate, stream0, valid0, mode, final_flag);
        timing_init(sm.g[1], s, t1, sb, so, dstate, stream1, valid1, mode, final_flag);
    } else if (role == 3) {
        afc_init(sm.g[0], s, a0, dstate, stream0);
        afc_init(sm.g[1], s, a1, dstate, stream1);
    }
    __syncthreads();
    if (role == 2) {  // prime both rings: everything up to w0 + 208..
        if (t0.live) while (t0.fill + kStageAll <= t0.w0 + kRingRows) stage_async(sm.g[0], s, t0.row, t0.stride, t0.fill, t0.w0);
        if (t1.live) while (t1.fill + kStageAll <= t1.w0 + kRingRows) stage_async(sm.g[1], s, t1.row, t1.stride, t1.fill, t1.w0);
        cp_async_commit();
        cp_async_wait<0>();
    }
    __syncthreads();

    for (;;) {
        if (!sm.g[0].any_live && !sm.g[1].any_live && !sm.g[0].ran && !sm.g[1].ran) break;  // uniform
        // ---- period A: window of group 0, loop of group 1
        if (role < 2) window_role(sm.g[0], s, role);
        else if (role == 2) timing_role(sm.g[1], s, t1, row0, mode, final_flag);
        else afc_role(sm.g[1], s, a1, afc_alpha);
        __syncthreads();
        // ---- period B: window of group 1, loop of group 0
        if (role < 2) window_role(sm.g[1], s, role);
        else if (role == 2) timing_role(sm.g[0], s, t0, row0, mode, final_flag);
        else afc_role(sm.g[0], s, a0, afc_alpha);
        __syncthreads();
    }

    // ---- persist: the timing warp writes the records, the AFC warp then patches its fields
    if (role == 2) {
        cp_async_wait<0>();
        timing_finish(t0, dstate, stream0, counters);
        timing_finish(t1, dstate, stream1, counters);
    }
    __syncthreads();
    if (role == 3) {
        if (valid0) afc_finish(a0, dstate, stream0);
        if (valid1) afc_finish(a1, dstate, stream1);
    }
}

cudaError_t launch_demod_pipe(const StreamBuffers& sb, const SoftBuffers& so, DemodState* dstate, int n_streams,
                              int mode, int final_flag, double afc_alpha, unsigned long long* counters,
                              cudaStream_t st) {
    const size_t smem = sizeof(PipeSmem);
    cudaError_t e = cudaFuncSetAttribute(demod_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int grid = (n_streams + 2 * kSpc - 1) / (2 * kSpc);
    demod_pipe_kernel<<<grid, kThreads, smem, st>>>(sb, so, dstate, n_streams, mode, final_flag, afc_alpha, counters);
    return cudaGetLastError();
}

}  // namespace opvd
