// kernels_demod_batch.cu — A1/A3/A4 for sm_100a, BATCHED variant for large channel banks
// (demod_batch_core.cuh).  Used when there are enough streams to fill the machine (thousands):
// the goal is throughput per issued instruction, not per-symbol latency (kernels_demod_warp.cu).
//
// A CTA of 128 threads (4 warps) owns 32 streams; lane = stream in every warp, so no instruction ever
// needs an intra-warp exchange and each one does 32 streams' worth of work.  Every symbol has two
// phases separated by CTA barriers:
//   window phase  warp k: tone k >> 1, window half k & 1.  30 samples -> three 10-sample Horner block
//                 sums -> three partial gates, interpolated (the post-sum interpolator is linear, so
//                 it is applied per half) -> shared memory.
//   loop phase    the recurrence, split into its two independent chains:
//                 warp 0 (TIMING chain): gate energies, soft decision, early-late TED, timing loop,
//                         next position, call schedule, soft-symbol store;
//                 warp 1 (AFC chain): on-time sums, phase detector (branch-free atan2), AFC loop,
//                         previous correlations, LO steps of the next symbol;
//                 each combines the helpers' partial gates it needs itself (cheaper than an exchange);
//                 warps 2-3 meanwhile move the next samples from HBM into the shared-memory ring.
// With one lane per stream the serial arithmetic of the recurrence is amortised over 32 streams (the
// warp-per-stream kernel spends 330 warp-instructions per stream and symbol, this kernel ~65 in
// total), and a 16,384-stream bank puts 14 warps on every SM.
//
// Sample ring: transposed, ring[row][stream] with row = sample index mod 256, so that lane s always
// reads bank s whatever its stream's window position is (per-stream rings laid out stream-major
// give 3-4-way bank conflicts on every load because the window offsets of the 32 streams are
// unrelated).  Rows 0..63 are mirrored behind row 255: a 61-row window never wraps.  Warps 2-3 fill
// it in the loop phase with 256-bit global loads (each thread 96 contiguous bytes of its stream per
// symbol, L2-prefetched at the top of the symbol and stored in the same loop phase: loads carried over
// the window phase stalled these warps on their shared register scoreboards), so HBM is read exactly
// once, in whole 32-byte sectors.
#include <cuda_runtime.h>
#include <cstdint>

#include "demod_batch_core.cuh"
#include "demod_warp_core.cuh"  // first_symbol_fix_w
#include "opvd_kernels.cuh"

namespace opvd {

namespace {

constexpr int kSpc = 32;            // streams per CTA
constexpr int kThreads = 128;       // 4 warps
constexpr int kRingRows = 256;      // samples per stream resident in shared memory (power of two)
constexpr int kMirrorRows = 64;     // rows 0..63 repeated after row 255
constexpr int kRows = kRingRows + kMirrorRows;
constexpr int kSub = 8;             // samples per 32-byte sector (2 x LDG.128)
constexpr int kStage = 3 * kSub;    // samples per staging thread and symbol
constexpr int kStageAll = 2 * kStage;  // per stream and symbol (two staging warps)

struct __align__(16) BatchSmem {
    uint32_t ring[kRows][kSpc];   // 40 KB
    double2 part[4][3][kSpc];     // [warp][E,O,L][stream] interpolated partial gates, 6 KB
    double zq[2][2][4][kSpc];     // [symbol parity][tone][z.r, z.i, q.r, q.i][stream], 4 KB: the AFC warp writes
                                  // the next symbol's LO steps while the timing warp still reads this symbol's
    double frac[kSpc];            // interpolation fraction f = pos - floor(pos) of the current symbol
    int w0[kSpc];                 // row-relative sample index of window slot 0 of the current symbol
    int live[kSpc];               // stream has a symbol to demodulate
    int first[kSpc];              // the current symbol is the first of a demodulate() call
    int any_live;
};

constexpr uint32_t kQBias = 0x80000000u;
// ring word (I raw, Q offset-binary) -> doubles.  I through I2F.F64.S16 (XU pipe), Q through the 2^52
// bias trick (integer pipe + one DADD): the window phase converts 60 components per thread and symbol
// and would otherwise be bound by the XU pipe (8 cycles per warp instruction).
__device__ __forceinline__ void unpack_ring(uint32_t w, double& I, double& Q) {
    I = (double)(int16_t)(w & 0xFFFFu);
    Q = __hiloint2double(0x43300000, (int)(w >> 16)) - 4503599627403264.0;  // 2^52 + 2^15
}
__device__ __forceinline__ uint32_t ring_raw(uint32_t w) { return w ^ kQBias; }

// 8 consecutive samples of stream s starting at sample index idx (multiple of 8)
__device__ __forceinline__ void stage_store8(BatchSmem& sm, int s, int idx, uint4 a, uint4 b) {
    const int row = idx & (kRingRows - 1);
    // the ring holds Q with its sign bit flipped (offset binary): unpack_ring() then builds the double
    // 2^52 + (Q + 32768) without a per-use XOR
    const uint32_t w[8] = {a.x ^ kQBias, a.y ^ kQBias, a.z ^ kQBias, a.w ^ kQBias,
                           b.x ^ kQBias, b.y ^ kQBias, b.z ^ kQBias, b.w ^ kQBias};
#pragma unroll
    for (int j = 0; j < 8; ++j) sm.ring[row + j][s] = w[j];
    if (row < kMirrorRows) {
#pragma unroll
        for (int j = 0; j < 8; ++j) sm.ring[kRingRows + row + j][s] = w[j];
    }
}
__device__ __forceinline__ void stage_store(BatchSmem& sm, int s, int idx, const uint4 (&v)[6]) {
#pragma unroll
    for (int c = 0; c < 3; ++c) stage_store8(sm, s, idx + kSub * c, v[2 * c], v[2 * c + 1]);
}
// 24 samples of a row starting at idx (multiple of 8); rows are 16-byte aligned and a multiple of 4
// samples long, so only the last chunk of a row can be partial
__device__ __forceinline__ void ldg256_stream(const uint32_t* p, uint4& a, uint4& b) {  // read-once, 32-byte aligned
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
                 : "l"(p));
}
// `wide`: rows are 32-byte aligned (row pointer and stride), so a whole sector travels in one 256-bit load: every
// lane reads its own line, and the LSU data pipe spends one wavefront per line and instruction, not per byte
__device__ __forceinline__ void stage_load(const uint32_t* row, int idx, int stride, bool wide, uint4 (&v)[6]) {
    const uint4* p = reinterpret_cast<const uint4*>(row + idx);
    if (idx + kStage <= stride) {
        if (wide) {
#pragma unroll
            for (int j = 0; j < 3; ++j) ldg256_stream(row + idx + 8 * j, v[2 * j], v[2 * j + 1]);
        } else {
#pragma unroll
            for (int j = 0; j < 6; ++j) v[j] = __ldg(p + j);
        }
    } else {
#pragma unroll
        for (int j = 0; j < 6; ++j) v[j] = (idx + 4 * j + 4 <= stride) ? __ldg(p + j) : make_uint4(0u, 0u, 0u, 0u);
    }
}

// early-gate correction for the first symbol of a call (rare: kept out of line)
__device__ __noinline__ cplx first_fix_cold(const uint32_t* win, double f, cplx z) {
    return first_symbol_fix_w([&](int kk) { return ring_raw(win[kk * kSpc]); }, f, z);
}
// call scheduling out of line.  Everything it touches by reference lives in local memory, so the
// hot loop hands it copies: the position comes back through st.pos.
__device__ __noinline__ bool schedule_cold(DemodState& st, int mode, long long avail, bool final_flag) {
    double pos = st.pos;
    const bool live = demod_schedule(st, pos, mode, avail, final_flag);
    st.pos = pos;
    return live;
}

__device__ __forceinline__ void publish_lo(BatchSmem& sm, int par, int s, const ToneLo& t1, const ToneLo& t2) {
    sm.zq[par][0][0][s] = t1.z.r; sm.zq[par][0][1][s] = t1.z.i; sm.zq[par][0][2][s] = t1.q.r; sm.zq[par][0][3][s] = t1.q.i;
    sm.zq[par][1][0][s] = t2.z.r; sm.zq[par][1][1][s] = t2.z.i; sm.zq[par][1][2][s] = t2.q.r; sm.zq[par][1][3][s] = t2.q.i;
}

// Loop phase, combining the two halves of one tone.  The two loop warps need different things (the
// timing chain the three gate energies of both tones, the AFC chain the on-time sums), and fetching
// them redundantly from the helpers' partial gates is cheaper than an exchange plus a barrier.
struct GateEnergies {
    double eE, eO, eL;
};
__device__ __forceinline__ GateEnergies finish_energies(BatchSmem& sm, int par, int s, int tone, int w0, double f,
                                                       bool first) {
    HalfGates a, b;
    double2 v;
    v = sm.part[2 * tone][0][s]; a.E = {v.x, v.y}; v = sm.part[2 * tone][1][s]; a.O = {v.x, v.y};
    v = sm.part[2 * tone][2][s]; a.L = {v.x, v.y};
    v = sm.part[2 * tone + 1][0][s]; b.E = {v.x, v.y}; v = sm.part[2 * tone + 1][1][s]; b.O = {v.x, v.y};
    v = sm.part[2 * tone + 1][2][s]; b.L = {v.x, v.y};
    ToneLo t;
    t.z = {sm.zq[par][tone][0][s], sm.zq[par][tone][1][s]};
    t.q = {sm.zq[par][tone][2][s], sm.zq[par][tone][3][s]};
    t.inc = 0.0;
    cplx fix = {0.0, 0.0};
    if (first)  // early-gate clamp (:237), once per call; window n is still in the ring
        fix = first_fix_cold(&sm.ring[w0 & (kRingRows - 1)][s], f, t.z);
    const ToneGates g = batch_finish_tone(a, b, t, fix);
    return {g.eE, g.eO, g.eL};
}
struct OnTime {
    cplx O, z40;
    double eO;
};
__device__ __forceinline__ OnTime finish_ontime(BatchSmem& sm, int par, int s, int tone) {
    double2 v;
    v = sm.part[2 * tone][1][s];
    const cplx aO = {v.x, v.y};
    v = sm.part[2 * tone + 1][1][s];
    const cplx bO = {v.x, v.y};
    const cplx q = {sm.zq[par][tone][2][s], sm.zq[par][tone][3][s]};
    const cplx q2 = csqr(q);
    OnTime o;
    o.O = cfma(q2, bO, aO);
    o.z40 = csqr(q2);
    o.eO = cnorm(o.O);
    return o;
}

}  // namespace

__global__ void __launch_bounds__(kThreads, 4)
demod_batch_kernel(StreamBuffers sb, SoftBuffers so, DemodState* __restrict__ dstate, int n_streams, int mode,
                   int final_flag, double afc_alpha, unsigned long long* __restrict__ counters) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BatchSmem& sm = *reinterpret_cast<BatchSmem*>(smem_raw);
    // roles rotate with the CTA index: warps land on SM sub-partitions by warp id, and the loop-phase
    // chains (roles 0 and 1) are heavier than the staging roles, so co-resident CTAs spread them out
    const int s = threadIdx.x & 31, k = ((threadIdx.x >> 5) + blockIdx.x) & 3;
    const int stream_raw = blockIdx.x * kSpc + s;
    const bool valid = stream_raw < n_streams;
    const int stream = valid ? stream_raw : n_streams - 1;
    const long long row0 = sb.row_base;
    const uint32_t* __restrict__ row = sb.iq + (long long)stream * sb.stride;  // row[r] = absolute sample row0 + r
    const int stride_i = (int)sb.stride;
    const bool wide = ((reinterpret_cast<uintptr_t>(sb.iq) | (uintptr_t)(sb.stride * 4)) & 31u) == 0;
    const bool prefetch = (long long)n_streams * sb.stride * 4 <= (48ll << 30);

    // ---- loop-phase state.  warp 0: timing chain + call schedule; warp 1: AFC chain
    DemodState st;                      // warp 0 (lives in local memory: only the scheduler touches it)
    double pos = 0.0, timing_freq = 0.0, call_len_d = 0.0;  // warp 0
    long long avail = 0, n_sym0 = 0, origin0 = 0;           // warp 0
    int origin_rel = 0, sym_in_call = 0;                    // warp 0
    double* soft_row = nullptr;
    double* soft_ptr = nullptr;                             // warp 0
    BatchAfc afc = {0.0, 0.0, 0.0, {0.0, 0.0}, {0.0, 0.0}};  // warp 1
    double inc1 = 0.0, inc2 = 0.0;                          // warp 1: LO phase steps of the current symbol
    if (k == 0) {
        st = dstate[stream];
        avail = sb.avail[stream];
        soft_row = so.soft + (long long)stream * so.stride - so.base;
        soft_ptr = soft_row + st.n_sym;
        n_sym0 = st.n_sym; origin0 = st.origin;
        timing_freq = st.timing_freq;
        const bool live = valid && schedule_cold(st, mode, avail, final_flag != 0);
        pos = st.pos;
        sym_in_call = st.sym_in_call;
        call_len_d = (double)st.call_len;
        origin_rel = (int)(st.origin - row0);
        int w0 = 0;
        double f = 0.0;
        if (live) {
            const int b = __double2int_rz(pos);  // pos >= 0: truncation == floor (:125)
            f = pos - (double)b;
            w0 = origin_rel + b - kWinLead;
        }
        sm.w0[s] = w0;
        sm.frac[s] = f;
        sm.live[s] = live ? 1 : 0;
        sm.first[s] = sym_in_call == 0;
        const int any = __any_sync(0xffffffffu, live);
        if (s == 0) sm.any_live = any;
    } else if (k == 1) {
        const DemodState* d = dstate + stream;
        afc.freq_offset = d->freq_offset; afc.ph1 = d->ph1; afc.ph2 = d->ph2; afc.p1 = d->p1; afc.p2 = d->p2;
        ToneLo t1, t2;
        batch_lo(afc.freq_offset, t1, t2);  // general version: a -o offset may exceed the fast range
        inc1 = t1.inc; inc2 = t2.inc;
        publish_lo(sm, 0, s, t1, t2);
    }
    __syncthreads();  // state of symbol 0 published

    // ---- prime the ring (warps 2-3): everything up to w0 + 208.. of each live stream
    int fill = 0;  // samples [.., fill) of this thread's stream have been requested (multiple of 8)
    if (k >= 2) {
        const int w0 = sm.w0[s];
        fill = (w0 < 0 ? 0 : w0) & ~(kSub - 1);
        if (sm.live[s]) {
            while (fill + kStageAll <= w0 + kRingRows) {
                const int idx = fill + kStage * (k - 2);
                if (idx < stride_i) {
                    uint4 v[6];
                    stage_load(row, idx, stride_i, wide, v);
                    stage_store(sm, s, idx, v);
                }
                fill += kStageAll;
            }
        }
    }
    __syncthreads();  // ring primed

    const int tone = k >> 1, half = k & 1;

    int par = 0;  // symbol parity: which half of zq holds the current symbol's LO steps
    while (sm.any_live) {
        // everything the loop phase overwrites for the next symbol is read here, before the barrier
        const int lv = sm.live[s];
        const int w0 = sm.w0[s];
        const bool first = sm.first[s] != 0;
        const double frac = sm.frac[s];
        // staging warps: pull the 96 bytes they will load in this symbol's loop phase into L2 now (no register and no
        // scoreboard is tied to a prefetch): the loop-phase loads then cost an L2 hit instead of an HBM round trip.
        // Only for banks whose captures span less than 48 GB: measured (demod kernel, Gsample/s, with / without the
        // prefetch) 4,096 streams x 6 frames 139 / 114, 8,192 x 6 256 / 214, 4,096 x 28 (40 GB) 131 / 109,
        // 18,944 x 6 (40 GB) 348 / 348, but 8,192 x 28 (80 GB) 174 / 206 and 18,944 x 14 (93 GB) 320 / 348 — beyond the
        // reach of the TLBs the hints are dropped and only cost LSU slots and page walks.
        if (prefetch && k >= 2 && lv && fill + kStageAll <= w0 + kRingRows) {
            const int idx = fill + kStage * (k - 2);
            if (idx + kStage <= stride_i) {
#pragma unroll
                for (int j = 0; j < 3; ++j) asm volatile("prefetch.global.L2 [%0];" ::"l"(row + idx + 8 * j));
            }
        }
        // ---- window phase (all four warps)
        if (lv) {
            const uint32_t* src = &sm.ring[(w0 & (kRingRows - 1)) + 30 * half][s];
            double I[31], Q[31];  // slots 30h .. 30h+29, and slot 60 for the late gate's edge term (h = 1)
#pragma unroll
            for (int j = 0; j < 30; ++j) unpack_ring(src[j * kSpc], I[j], Q[j]);
            I[30] = 0.0; Q[30] = 0.0;
            if (half) unpack_ring(src[30 * kSpc], I[30], Q[30]);
            const cplx z = {sm.zq[par][tone][0][s], sm.zq[par][tone][1][s]};
            const cplx q = {sm.zq[par][tone][2][s], sm.zq[par][tone][3][s]};
            const HalfGates g = batch_half_gates(I, Q, z, q, frac, half);
            sm.part[k][0][s] = make_double2(g.E.r, g.E.i);
            sm.part[k][1][s] = make_double2(g.O.r, g.O.i);
            sm.part[k][2][s] = make_double2(g.L.r, g.L.i);
        }
        __syncthreads();  // partial gates ready
        if (k == 0) {
            // ---- gate energies of both tones, then the timing chain
            bool live = lv != 0;
            if (live) {
                const GateEnergies g1 = finish_energies(sm, par, s, 0, w0, frac, first);
                const GateEnergies g2 = finish_energies(sm, par, s, 1, w0, frac, first);
                const double soft = batch_timing(g1.eO, g2.eO, g1.eE, g1.eL, g2.eE, g2.eL, timing_freq, pos, g_fm);
                *soft_ptr++ = soft;
                sym_in_call = 1;  // any non-zero value: the open call has produced symbols
                // ---- next symbol of this stream
                if (!((pos + 40.0) + 10.0 < call_len_d)) {  // :221 fails: close the call, maybe open the next
                    st.n_sym = (long long)(soft_ptr - soft_row);
                    st.sym_in_call = sym_in_call;
                    st.pos = pos;
                    live = schedule_cold(st, mode, avail, final_flag != 0);
                    pos = st.pos;
                    sym_in_call = st.sym_in_call;
                    call_len_d = (double)st.call_len;
                    origin_rel = (int)(st.origin - row0);
                }
                if (live) {
                    const int b2 = __double2int_rz(pos);
                    sm.w0[s] = origin_rel + b2 - kWinLead;
                    sm.frac[s] = pos - (double)b2;
                    sm.first[s] = sym_in_call == 0;
                } else {
                    sm.live[s] = 0;
                }
            }
            const int any = __any_sync(0xffffffffu, live);
            if (s == 0) sm.any_live = any;
        } else if (k == 1) {
            // ---- on-time sums of both tones, then the AFC chain and the LO steps of the next symbol
            if (lv) {
                const OnTime o1 = finish_ontime(sm, par, s, 0), o2 = finish_ontime(sm, par, s, 1);
                batch_afc(afc, o1.O, o1.z40, o1.eO, o2.O, o2.z40, o2.eO, inc1, inc2, first, afc_alpha, g_fm);
                if (!first) {
                    ToneLo t1, t2;
                    batch_lo_fast(afc.freq_offset, t1, t2, g_fm);  // |freq_offset| <= 2 kHz after the AFC clamp
                    inc1 = t1.inc; inc2 = t2.inc;
                    publish_lo(sm, par ^ 1, s, t1, t2);
                } else {  // no AFC update on the first symbol of a call (:289): same LO steps next symbol
#pragma unroll
                    for (int t = 0; t < 2; ++t)
#pragma unroll
                        for (int c = 0; c < 4; ++c) sm.zq[par ^ 1][t][c][s] = sm.zq[par][t][c][s];
                }
            }
        } else {
            // ---- staging (warps 2-3): request the next 24 samples per stream and store them in the SAME loop phase (their
            // rows hold samples older than the window that has just been read).  The loads are not carried over the
            // window phase: a warp with global loads in flight shares its register scoreboards with them, and the
            // staging warps used to stall at the top of the window phase until HBM answered (8 % of all warp samples,
            // `@!P0 BRA` long_scoreboard in the source-level profile), holding the whole CTA at the next barrier.  Here
            // they wait while the timing and AFC warps run their chains, which take longer than an HBM round trip.
            if (lv && fill + kStageAll <= w0 + kRingRows) {
                const int idx = fill + kStage * (k - 2);
                if (idx < stride_i) {
                    uint4 pend[6];
                    stage_load(row, idx, stride_i, wide, pend);
                    stage_store(sm, s, idx, pend);
                }
                fill += kStageAll;
            }
        }
        __syncthreads();  // state of the next symbol published, ring advanced
        par ^= 1;
    }

    // ---- persist the streams' state: warp 0 writes the record, warp 1 then patches the AFC fields
    if (k == 0 && valid) {
        st.n_sym = (long long)(soft_ptr - soft_row);
        st.sym_in_call = sym_in_call;
        st.pos = pos; st.timing_freq = timing_freq;
        dstate[stream] = st;
        unsigned long long dsym = (unsigned long long)(st.n_sym - n_sym0);
        unsigned long long dsmp = (unsigned long long)(st.origin - origin0);
        if (st.flags & kFlagDone) dsmp = (unsigned long long)(avail - origin0);
        if (dsym) atomicAdd(&counters[kCtrSymbols], dsym);
        if (dsmp) atomicAdd(&counters[kCtrSamples], dsmp);
    }
    __syncthreads();
    if (k == 1 && valid) {
        DemodState* d = dstate + stream;
        d->freq_offset = afc.freq_offset; d->ph1 = afc.ph1; d->ph2 = afc.ph2; d->p1 = afc.p1; d->p2 = afc.p2;
    }
}

cudaError_t launch_demod_batch(const StreamBuffers& sb, const SoftBuffers& so, DemodState* dstate, int n_streams,
                               int mode, int final_flag, double afc_alpha, unsigned long long* counters,
                               cudaStream_t st) {
    const size_t smem = sizeof(BatchSmem);
    cudaError_t e = cudaFuncSetAttribute(demod_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int grid = (n_streams + kSpc - 1) / kSpc;
    demod_batch_kernel<<<grid, kThreads, smem, st>>>(sb, so, dstate, n_streams, mode, final_flag, afc_alpha, counters);
    return cudaGetLastError();
}

}  // namespace opvd
