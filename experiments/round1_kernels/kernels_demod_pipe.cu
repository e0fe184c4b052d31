// kernels_demod_pipe.cu — A1/A3/A4 for sm_100a, PIPELINED batched variant for large channel banks.
// Same arithmetic as kernels_demod_batch.cu (demod_batch_core.cuh); different schedule.
//
// The batched kernel alternates a window phase and a loop phase per symbol, so at any time about half
// of its warps wait at a barrier.  Here a CTA of 512 threads (one per SM) is two independent QUADS of eight
// warps; a quad owns TWO groups of 32 streams (lane = stream everywhere) and its warps are specialised:
//   4 window workers (tone t, half h): 30 samples of the 60-sample window for one tone in two passes (blocks
//                 0-1, then block 2: the register budget at 512 threads is 128), 10-sample block sums by two
//                 5-step Horner chains joined by z^5, interpolated partial gates; after a two-warp named
//                 barrier the half-0 worker of the tone combines the halves and publishes the tone's gate
//                 energies / on-time sum;
//   timing warp   timing chain of the group whose window finished one period earlier (soft decision, TED,
//                 timing loop, call schedule, soft store);
//   AFC warp      AFC chain of the same group (phase detector, AFC loop, LO steps z, z^5, z^10 of its next symbol);
//   2 staging warps, one per group: request 48 samples per stream in the group's window period (coalesced:
//                 the 16-byte pieces are dealt to the lanes in memory order), move them through a landing
//                 buffer into the transposed shared-memory ring in the group's loop period.
// Period p: window(group p & 1) runs while loop(group ~p & 1) runs; one quad barrier (named, 256 threads) per
// period.  Warps land on SM sub-partitions by warp id: every sub-partition hosts one window worker of each
// quad, and the other roles are rotated between the quads.  DESIGN.md section 3.2.1 lists the measurements
// behind each of these choices.
#include <cuda_runtime.h>
#include <cstdint>

#include "demod_batch_core.cuh"
#include "demod_warp_core.cuh"  // first_symbol_fix_w
#include "opvd_kernels.cuh"

namespace opvd {

namespace {

constexpr int kSpc = 32;            // streams per group
constexpr int kQuads = 2;           // quads per CTA
constexpr int kGroups = 2 * kQuads; // stream groups per CTA
constexpr int kQuadWarps = 8;      // four window workers (tone x half), timing, AFC, two staging warps
constexpr int kThreads = 32 * kQuadWarps * kQuads;
constexpr int kRingRows = 256;      // samples per stream resident in shared memory (power of two)
constexpr int kMirrorRows = 64;     // rows 0..63 are repeated after row 255 (the first 60 of them are stored):
constexpr int kRows = kRingRows + kWin - 1;  // a 61-row window starting at row <= 255 never wraps
constexpr int kSub = 8;             // samples per 32-byte sector (2 x LDG.128)
constexpr int kStageVec = 12;       // uint4 per stream and visit
constexpr int kStageAll = 4 * kStageVec;  // samples staged per stream and symbol (48)

struct __align__(16) GroupSmem {
    uint32_t ring[kRows][kSpc];     // transposed sample ring (Q offset-binary), 40 KB
    double2 part[2][2][3][kSpc];    // [tone][half][E,O,L] interpolated partial gates
    double tg[2][7][kSpc];          // per tone: eE, eO, eL, O.r, O.i, z40.r, z40.i
    double zq[2][6][kSpc];          // [tone][z.r, z.i, q.r, q.i, z5.r, z5.i]: LO steps of the group's next window
    double frac[kSpc];              // interpolation fraction of the next window
    int w0[kSpc];                   // row-relative sample index of slot 0 of the next window
    int live[kSpc], first[kSpc];    // next window: stream has a symbol / it is the first of a call
    int sym_live[kSpc], sym_first[kSpc];  // the same two for the symbol the loop warps are working on
    int any_live;                   // some stream of the group has a next window
    int ran;                        // the window workers processed the group in their last period
};
constexpr int kStagePitch = kStageVec + 1;  // uint4 per stream in the landing buffer: 208 B, so that the 8 lanes of
                                            // a quarter-warp read 128-bit words from 32 different banks
struct __align__(16) PipeSmem {
    GroupSmem g[kGroups];
    uint4 landing[kQuads][kSpc][kStagePitch];  // transposition buffer of a quad's staging warp (linear per stream)
};
static_assert(sizeof(PipeSmem) <= 227 * 1024, "pipe kernel shared memory");

// named barriers: 1 + 3*quad = the quad (256 threads), 2 + 3*quad + tone = the two window workers of a tone
__device__ __forceinline__ void quad_barrier(int id) { asm volatile("bar.sync %0, 256;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void pair_barrier(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

constexpr uint32_t kQBias = 0x80000000u;
// ring word (I raw, Q offset-binary) -> doubles.  I through I2F.F64.S16 (XU pipe), Q through the 2^52
// bias trick (integer pipe + one DADD on the FP64 pipe).
__device__ __forceinline__ void unpack_ring(uint32_t w, double& I, double& Q) {
    I = (double)(int16_t)(w & 0xFFFFu);
    Q = __hiloint2double(0x43300000, (int)(w >> 16)) - 4503599627403264.0;  // 2^52 + 2^15
}

__device__ __noinline__ cplx first_fix_cold(const uint32_t* win, double f, cplx z) {
    return first_symbol_fix_w([&](int kk) { return win[kk * kSpc] ^ kQBias; }, f, z);
}
__device__ __noinline__ bool schedule_cold(DemodState& st, int mode, long long avail, bool final_flag) {
    double pos = st.pos;
    const bool live = demod_schedule(st, pos, mode, avail, final_flag);
    st.pos = pos;
    return live;
}

// ---- ring staging (timing warp): 48 samples of one stream per visit, through registers
// rows are 16-byte aligned and a multiple of 4 samples long, so only the last chunk of a row can be partial
__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {  // read-once data: keep it out of the (tiny) L1
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void stage_load(const uint32_t* row, int idx, int stride, uint4 (&v)[kStageVec]) {
    const uint4* p = reinterpret_cast<const uint4*>(row + idx);
    if (idx + kStageAll <= stride) {
#pragma unroll
        for (int j = 0; j < kStageVec; ++j) v[j] = ldg_stream(p + j);
    } else {
#pragma unroll
        for (int j = 0; j < kStageVec; ++j) v[j] = (idx + 4 * j + 4 <= stride) ? ldg_stream(p + j) : make_uint4(0u, 0u, 0u, 0u);
    }
}
__device__ __forceinline__ void stage_store(GroupSmem& sm, int s, int idx, const uint4 (&v)[kStageVec]) {
#pragma unroll
    for (int c = 0; c < kStageVec / 2; ++c) {
        const int row = (idx + kSub * c) & (kRingRows - 1);
        const uint4 a = v[2 * c], b = v[2 * c + 1];
        // the ring holds Q with its sign bit flipped (offset binary): unpack_ring() needs no per-use XOR
        const uint32_t w[8] = {a.x ^ kQBias, a.y ^ kQBias, a.z ^ kQBias, a.w ^ kQBias,
                               b.x ^ kQBias, b.y ^ kQBias, b.z ^ kQBias, b.w ^ kQBias};
#pragma unroll
        for (int j = 0; j < 8; ++j) sm.ring[row + j][s] = w[j];
        if (row < kMirrorRows) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (j < 4 || row + 8 <= kRows - kRingRows) sm.ring[kRingRows + row + j][s] = w[j];  // rows 256..315
        }
    }
}

// ---- window worker (tone, half): 30 samples of the window for one tone; then the half-0 worker of the tone
// combines the two halves and publishes the tone's gate energies / on-time sum
__device__ __forceinline__ void window_role(GroupSmem& sm, int s, int tone, int half, int pbar) {
    if (!sm.any_live) {  // uniform
        if ((tone | half) == 0 && s == 0) sm.ran = 0;
        return;
    }
    const int lv = sm.live[s], w0 = sm.w0[s];
    const bool first = sm.first[s] != 0;
    const double f = sm.frac[s];
    if ((tone | half) == 0) {
        sm.sym_live[s] = lv;
        sm.sym_first[s] = first;
        if (s == 0) sm.ran = 1;
    }
    const uint32_t* win = &sm.ring[w0 & (kRingRows - 1)][s];
    const cplx z = {sm.zq[tone][0][s], sm.zq[tone][1][s]}, q = {sm.zq[tone][2][s], sm.zq[tone][3][s]};
    const cplx z5 = {sm.zq[tone][4][s], sm.zq[tone][5][s]};
    if (lv) {
        // two passes keep the register footprint small (a spill is an L2 round trip here: nearly all of L1 is
        // carved out as shared memory): blocks 0-1 of the half, then block 2
        const uint32_t* src = win + 30 * half * kSpc;
        cplx A, B, C, s0, s10, s20, s30;
        {
            double I[20], Q[20];
#pragma unroll
            for (int j = 0; j < 20; ++j) unpack_ring(src[j * kSpc], I[j], Q[j]);
            A = horner10_split(I, Q, z, z5); B = horner10_split(I + 10, Q + 10, z, z5);
            s0 = {I[0], Q[0]}; s10 = {I[10], Q[10]};
        }
        {
            double I[11], Q[11];  // slots 30h+20 .. 30h+29, and slot 60 for the late gate's edge term (h = 1)
#pragma unroll
            for (int j = 0; j < 10; ++j) unpack_ring(src[(20 + j) * kSpc], I[j], Q[j]);
            I[10] = 0.0; Q[10] = 0.0;
            if (half) unpack_ring(src[30 * kSpc], I[10], Q[10]);
            C = horner10_split(I, Q, z, z5);
            s20 = {I[0], Q[0]}; s30 = {I[10], Q[10]};
        }
        const HalfGates g = half_gates_from_blocks(A, B, C, half ? s10 : s0, half ? s20 : s10, half ? s30 : s20, z, q, f, half);
        sm.part[tone][half][0][s] = make_double2(g.E.r, g.E.i);
        sm.part[tone][half][1][s] = make_double2(g.O.r, g.O.i);
        sm.part[tone][half][2][s] = make_double2(g.L.r, g.L.i);
    }
    pair_barrier(pbar);
    if (lv && half == 0) {
        HalfGates a, b;
        double2 v;
        v = sm.part[tone][0][0][s]; a.E = {v.x, v.y}; v = sm.part[tone][0][1][s]; a.O = {v.x, v.y};
        v = sm.part[tone][0][2][s]; a.L = {v.x, v.y};
        v = sm.part[tone][1][0][s]; b.E = {v.x, v.y}; v = sm.part[tone][1][1][s]; b.O = {v.x, v.y};
        v = sm.part[tone][1][2][s]; b.L = {v.x, v.y};
        ToneLo t;
        t.z = z; t.q = q; t.inc = 0.0;
        cplx fix = {0.0, 0.0};
        if (first) fix = first_fix_cold(win, f, t.z);  // early-gate clamp (:237), once per call
        const ToneGates g = batch_finish_tone(a, b, t, fix);
        sm.tg[tone][0][s] = g.eE; sm.tg[tone][1][s] = g.eO; sm.tg[tone][2][s] = g.eL;
        sm.tg[tone][3][s] = g.O.r; sm.tg[tone][4][s] = g.O.i; sm.tg[tone][5][s] = g.z40.r; sm.tg[tone][6][s] = g.z40.i;
    }
}

// ---- timing warp state of one group
// (the DemodState record itself is a separate object: the out-of-line scheduler takes its address, which
// pins it in local memory, and with ~215 KB of shared memory carved out of L1 every local access is an
// L2 round trip — nothing the hot path touches may share an object with it)
struct TimingState {
    double pos, timing_freq, call_len_d;
    long long avail, n_sym0, origin0;
    double* soft_row;
    double* soft_ptr;
    int origin_rel, sym_in_call;
    bool live, valid;
};

__device__ __forceinline__ void timing_publish(GroupSmem& sm, int s, TimingState& t) {
    if (t.live) {
        const int b = __double2int_rz(t.pos);  // pos >= 0: truncation == floor (:125)
        sm.w0[s] = t.origin_rel + b - kWinLead;
        sm.frac[s] = t.pos - (double)b;
        sm.first[s] = t.sym_in_call == 0;
    }
    sm.live[s] = t.live ? 1 : 0;
    const int any = __any_sync(0xffffffffu, t.live);
    if (s == 0) sm.any_live = any;
}

__device__ __forceinline__ void timing_init(GroupSmem& sm, int s, TimingState& t, DemodState& st, const StreamBuffers& sb,
                                            const SoftBuffers& so, const DemodState* dstate, int stream, bool valid,
                                            int mode, int final_flag) {
    t.valid = valid;
    st = dstate[stream];
    t.avail = sb.avail[stream];
    t.soft_row = so.soft + (long long)stream * so.stride - so.base;
    t.soft_ptr = t.soft_row + st.n_sym;
    t.n_sym0 = st.n_sym; t.origin0 = st.origin;
    t.timing_freq = st.timing_freq;
    t.live = valid && schedule_cold(st, mode, t.avail, final_flag != 0);
    t.pos = st.pos;
    t.sym_in_call = st.sym_in_call;
    t.call_len_d = (double)st.call_len;
    t.origin_rel = (int)(st.origin - sb.row_base);
    sm.w0[s] = 0;
    timing_publish(sm, s, t);
    if (s == 0) sm.ran = 0;
}

__device__ __forceinline__ void timing_role(GroupSmem& sm, int s, TimingState& t, DemodState& st, long long row0, int mode,
                                            int final_flag) {
    if (!sm.ran) return;  // uniform: the group's window did not run, nothing to do
    const int sym_live = sm.sym_live[s];
    if (sym_live) {
        const double soft = batch_timing(sm.tg[0][1][s], sm.tg[1][1][s], sm.tg[0][0][s], sm.tg[0][2][s], sm.tg[1][0][s],
                                         sm.tg[1][2][s], t.timing_freq, t.pos, g_fm);
        *t.soft_ptr++ = soft;
        t.sym_in_call = 1;  // any non-zero value: the open call has produced symbols
        if (!((t.pos + 40.0) + 10.0 < t.call_len_d)) {  // :221 fails: close the call, maybe open the next
            st.n_sym = (long long)(t.soft_ptr - t.soft_row);
            st.sym_in_call = t.sym_in_call;
            st.pos = t.pos;
            t.live = schedule_cold(st, mode, t.avail, final_flag != 0);
            t.pos = st.pos;
            t.sym_in_call = st.sym_in_call;
            t.call_len_d = (double)st.call_len;
            t.origin_rel = (int)(st.origin - row0);
        }
    }
    timing_publish(sm, s, t);
}

// ---- staging warp of a quad: keeps the two groups' rings ahead of their windows.  In every period it first
// stores the 48 samples per stream it requested one period ago (for the group whose window ran then; the
// rows it overwrites hold samples older than that window, so nobody reads them any more), then requests
// the next 48 for the group whose window runs NOW (its w0/live are stable: the timing warp is on the other
// group).  The stores are complete before that group's next window starts, one barrier later.  A dedicated
// warp, because a warp with global loads in flight shares its register scoreboards with them: when the
// timing warp staged, its visit took a full HBM latency (measured).
struct StageState {
    const uint32_t* row;
    int fill;
};
typedef uint4 PendRegs[kStageVec];
__device__ __forceinline__ void stage_init(GroupSmem& sm, int s, StageState& g, PendRegs& pend, const StreamBuffers& sb,
                                           int stream) {
    g.row = sb.iq + (long long)stream * sb.stride;
    const int stride = (int)sb.stride, w0 = sm.w0[s];
    g.fill = (w0 < 0 ? 0 : w0) & ~(kSub - 1);
    if (sm.live[s]) {  // prime the ring: everything up to w0 + 208..
#pragma unroll 1
        while (g.fill + kStageAll <= w0 + kRingRows) {
            if (g.fill < stride) {
                stage_load(g.row, g.fill, stride, pend);
                stage_store(sm, s, g.fill, pend);
            }
            g.fill += kStageAll;
        }
    }
}

__device__ __forceinline__ void timing_finish(TimingState& t, DemodState& st, DemodState* dstate, int stream,
                                              unsigned long long* counters) {
    if (!t.valid) return;
    st.n_sym = (long long)(t.soft_ptr - t.soft_row);
    st.sym_in_call = t.sym_in_call;
    st.pos = t.pos; st.timing_freq = t.timing_freq;
    dstate[stream] = st;
    unsigned long long dsym = (unsigned long long)(st.n_sym - t.n_sym0);
    unsigned long long dsmp = (unsigned long long)(st.origin - t.origin0);
    if (st.flags & kFlagDone) dsmp = (unsigned long long)(t.avail - t.origin0);
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
    sm.zq[0][4][s] = t1.z5.r; sm.zq[0][5][s] = t1.z5.i; sm.zq[1][4][s] = t2.z5.r; sm.zq[1][5][s] = t2.z5.i;
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

template <typename T>
__device__ __forceinline__ void swap_regs(T& a, T& b) { const T t = a; a = b; b = t; }

}  // namespace

// Code size matters as much as the schedule here: the ten warps of a CTA run five different loops, and
// the SM's instruction cache (L1.5, 32 KB) must hold all of them — a version with one inlined copy of each
// role per group and quad (128 KB of SASS) spent 30-60 % of its issue slots waiting for instructions.  So
// every role has ONE copy of its loop body; the two groups alternate through a pointer swap (window and
// staging warps) plus a register swap of the per-group loop state (timing and AFC warps).
__global__ void __launch_bounds__(kThreads, 1)
demod_pipe_kernel(StreamBuffers sb, SoftBuffers so, DemodState* __restrict__ dstate, int n_streams, int mode,
                  int final_flag, double afc_alpha, unsigned long long* __restrict__ counters) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PipeSmem& sm = *reinterpret_cast<PipeSmem*>(smem_raw);
    const int s = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // Warps land on SM sub-partitions by warp id mod 4.  In each quad warps 0-3 are the window workers (tone =
    // w >> 1, half = w & 1), so every sub-partition hosts one window worker of each quad.  The other four roles
    // (4 timing, 5 AFC, 6 / 7 staging of the quad's first / second group) are rotated by two warps in quad 1, so
    // that the two AFC warps (as much FP64 work as a window worker) sit on different sub-partitions.
    const int quad = warp / kQuadWarps;
    const int wq = warp - kQuadWarps * quad;
    const int role = wq < 4 ? wq : 4 + ((wq + 2 * quad) & 3);
    const int qbar = 1 + 3 * quad, pbar = 2 + 3 * quad + ((role >> 1) & 1);
    GroupSmem* gw = &sm.g[2 * quad];  // group whose WINDOW runs in the current period (period 0: the quad's first)
    GroupSmem* gl = gw + 1;           // group whose LOOP runs in the current period
    const int raw0 = (blockIdx.x * kGroups + 2 * quad) * kSpc + s, raw1 = raw0 + kSpc;
    const bool valid0 = raw0 < n_streams, valid1 = raw1 < n_streams;
    const int stream0 = valid0 ? raw0 : n_streams - 1, stream1 = valid1 ? raw1 : n_streams - 1;
    const long long row0 = sb.row_base;

    // Period p: window of gw, loop of gl, then the two swap.  The exit test at the top of a period reads
    // only flags written during the PREVIOUS period and not rewritten in this one (gl->ran by W0,
    // gw->any_live by T), so a warp that is already inside the period cannot change what a slower warp
    // still has to read.  It is complete: !gl->ran means gl's window found gl->any_live == 0, which T
    // (skipping gl from now on) never sets again, and gw->any_live == 0 was written by the visit of T that
    // consumed gw's last window.  Periods 0 and 1 start the pipeline and are not tested.
#define OPVD_PIPE_EXIT(p) ((p) >= 2 && !gl->ran && !gw->any_live)

    if (role < 4) {
        quad_barrier(qbar);  // loop state of symbol 0 published
        quad_barrier(qbar);  // rings primed
#pragma unroll 1
        for (int p = 0;; ++p) {
            if (OPVD_PIPE_EXIT(p)) break;
            window_role(*gw, s, role >> 1, role & 1, pbar);
            quad_barrier(qbar);
            swap_regs(gw, gl);
        }
        quad_barrier(qbar);
    } else if (role == 4) {
        TimingState tw, tl;   // state of gw's / gl's streams
        DemodState st0, st1;  // local memory: only the out-of-line scheduler touches them
        DemodState* stw = &st0;
        DemodState* stl = &st1;
        int strw = stream0, strl = stream1;
        timing_init(*gw, s, tw, st0, sb, so, dstate, stream0, valid0, mode, final_flag);
        timing_init(*gl, s, tl, st1, sb, so, dstate, stream1, valid1, mode, final_flag);
        quad_barrier(qbar);
        quad_barrier(qbar);
#pragma unroll 1
        for (int p = 0;; ++p) {
            if (OPVD_PIPE_EXIT(p)) break;
            timing_role(*gl, s, tl, *stl, row0, mode, final_flag);
            quad_barrier(qbar);
            swap_regs(gw, gl); swap_regs(tw, tl); swap_regs(stw, stl); swap_regs(strw, strl);
        }
        // ---- persist: the timing warp writes the records, the AFC warp then patches its fields
        timing_finish(tw, *stw, dstate, strw, counters);
        timing_finish(tl, *stl, dstate, strl, counters);
        quad_barrier(qbar);
    } else if (role == 5) {
        AfcState aw, al;
        afc_init(*gw, s, aw, dstate, stream0);
        afc_init(*gl, s, al, dstate, stream1);
        bool w_is_0 = true;
        quad_barrier(qbar);
        quad_barrier(qbar);
#pragma unroll 1
        for (int p = 0;; ++p) {
            if (OPVD_PIPE_EXIT(p)) break;
            afc_role(*gl, s, al, afc_alpha);
            quad_barrier(qbar);
            swap_regs(gw, gl); swap_regs(aw, al); w_is_0 = !w_is_0;
        }
        quad_barrier(qbar);
        const AfcState& a0 = w_is_0 ? aw : al;
        const AfcState& a1 = w_is_0 ? al : aw;
        if (valid0) afc_finish(a0, dstate, stream0);
        if (valid1) afc_finish(a1, dstate, stream1);
    } else {
        // ---- staging warp of ONE group (role 6: the quad's first, role 7: its second).  In the group's window
        // period it requests the next 48 samples per stream (w0/live are stable then: the timing warp is on the
        // other group); in the group's loop period (the loads have had a whole period to land) it moves them into
        // the transposed ring.  The rows it overwrites hold samples older than the window that ran when they were
        // requested, and the stores are complete one barrier before the group's next window.
        // Load layout: the 384 16-byte pieces of a round are dealt to the lanes in memory order (piece m = 32 i +
        // lane of instruction i belongs to stream m / 12), so one instruction touches ~6 lines instead of 32; the
        // pieces then go through the linear landing buffer (128-bit stores and loads, conflict-free at a 208-byte
        // pitch) to the lane that owns the stream.  With one line per lane the loads took 32 wavefronts each on the
        // LSU data pipe and delayed every other warp's shared-memory loads; per-lane bulk copies (UBLKCP) serialise
        // at ~75 cycles per copy (both measured).
        const bool second = role == 7;
        GroupSmem* mine = second ? gl : gw;
        StageState f;
        PendRegs pend;
        const int stride = (int)sb.stride;
        uint4 (*land)[kStagePitch] = sm.landing[quad];  // shared by the quad's two staging warps: they store in
                                                        // alternate periods
        const uint32_t* base = sb.iq + (long long)((second ? raw1 : raw0) - s) * sb.stride;  // the group's first stream
        quad_barrier(qbar);
        stage_init(*mine, s, f, pend, sb, second ? stream1 : stream0);
        quad_barrier(qbar);
        int land_idx = -1;       // >= 0: pend holds pieces of samples [land_idx, land_idx + 48) of this lane's stream
        bool in_flight = false;  // uniform: pend holds a round
#pragma unroll 1
        for (int p = 0;; ++p) {
            if (OPVD_PIPE_EXIT(p)) break;
            if (gw == mine) {  // the group's window period: request
                land_idx = -1;
                if (mine->any_live) {  // uniform
                    if (mine->live[s] && f.fill + kStageAll <= mine->w0[s] + kRingRows) {
                        if (f.fill < stride) land_idx = f.fill;
                        f.fill += kStageAll;
                    }
                    if (__any_sync(0xffffffffu, land_idx >= 0)) {
#pragma unroll
                        for (int i = 0; i < kStageVec; ++i) {
                            const int m = 32 * i + s, sig = m / kStageVec, c = m - kStageVec * sig;
                            const int ridx = __shfl_sync(0xffffffffu, land_idx, sig);
                            uint4 v = make_uint4(0u, 0u, 0u, 0u);
                            if (ridx >= 0 && ridx + 4 * c + 4 <= stride)
                                v = ldg_stream(reinterpret_cast<const uint4*>(base + (long long)sig * stride + ridx) + c);
                            pend[i] = v;
                        }
                        in_flight = true;
                    }
                }
            } else if (in_flight) {  // the group's loop period: store
#pragma unroll
                for (int i = 0; i < kStageVec; ++i) {
                    const int m = 32 * i + s, sig = m / kStageVec;
                    land[sig][m - kStageVec * sig] = pend[i];
                }
                __syncwarp();
                if (land_idx >= 0) {
#pragma unroll
                    for (int j = 0; j < kStageVec; ++j) pend[j] = land[s][j];
                    stage_store(*mine, s, land_idx, pend);
                }
                in_flight = false;
            }
            quad_barrier(qbar);
            swap_regs(gw, gl);
        }
        quad_barrier(qbar);
    }
#undef OPVD_PIPE_EXIT
}

cudaError_t launch_demod_pipe(const StreamBuffers& sb, const SoftBuffers& so, DemodState* dstate, int n_streams,
                              int mode, int final_flag, double afc_alpha, unsigned long long* counters,
                              cudaStream_t st) {
    const size_t smem = sizeof(PipeSmem);
    cudaError_t e = cudaFuncSetAttribute(demod_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int per_cta = kGroups * kSpc;
    const int grid = (n_streams + per_cta - 1) / per_cta;
    demod_pipe_kernel<<<grid, kThreads, smem, st>>>(sb, so, dstate, n_streams, mode, final_flag, afc_alpha, counters);
    return cudaGetLastError();
}

}  // namespace opvd
