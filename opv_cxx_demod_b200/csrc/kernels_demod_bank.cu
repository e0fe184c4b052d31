// kernels_demod_bank.cu — A1/A3/A4 for sm_100a, CHANNEL-BANK variant (demod_bank_core.cuh): the kernel for
// thousands of streams (north-star regime, >= 16,384 streams per GPU = 128 per SM).
//
// A CTA of 96 threads owns 32 streams; lane = stream in each of its three warps, so every instruction does 32
// streams' worth of work and nothing is ever exchanged inside a warp.  The warps are free-running ROLES coupled only
// by named barriers (producer/consumer hand-offs through shared memory), never by a CTA barrier:
//   WINDOW  Horner block sums -> gate combination -> soft symbol, dominant tone -> early/late gates of the
//           dominant tone only -> TED, timing loop, next position, call schedule (:221-286, :313, :1012-1113)
//   AFC     phase detector, AFC loop, LO steps z (handed back first: the next symbol's Horner needs nothing else),
//           then the LO powers z^10, z^20, zeta^40 for the gate combination (:289-310)
//   STAGE   HBM -> transposed shared-memory ring, one batch of 32-byte sectors per lane in flight, 2-4 symbols
//           ahead of the window; paces itself with __nanosleep (about one round per symbol)
// The AFC chain of symbol n overlaps the early/late + timing part of symbol n.  Hardware placement (tools/warp_place.cu):
// warp w of the j-th resident 3-warp CTA sits on SM sub-partition (3 j + w) & 3, so with four resident CTAs every
// sub-partition hosts exactly one warp of each role.
//
// Two other organisations were built and measured in round 2 and lost (experiments/round2_kernels/):
// cutting the window over two dependent warps (22.7 ms on the 18,944 x 6 probe bank) and 16 streams x 2 window halves
// per warp with shuffles (25.4 ms), against 19.6 ms for this one.
// Sample ring: ring[row][stream], row = sample index mod 256, rows 0..63 mirrored behind row 255 so a 61-row
// window never wraps; lane s always reads bank s (conflict-free whatever the streams' window positions are).
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdlib>

#include "demod_bank_core.cuh"
#include "demod_warp_core.cuh"  // first_symbol_fix_w
#include "opvd_kernels.cuh"

namespace opvd {

namespace {

constexpr int kSpc = 32;             // streams per CTA
constexpr int kMirrorRows = 64;      // rows 0..63 of the ring repeated behind its last row
constexpr int kChunk = 8;            // samples per 32-byte sector
constexpr unsigned kFull = 0xffffffffu;
constexpr uint32_t kQBias = 0x80000000u;

__device__ unsigned g_stage_sleep_ns = 400;  // pacing of the STAGE warp; flat between 300 and 1,000 ns (19.85 .. 20.0 ms on the full
                                              // probe bank, profiles/gpu_r02_sleep.log).  OPVD_BANK_SLEEP: development switch

// named barriers.  The two CTA-wide rendezvous (after set-up, before exit) are reached from three different role
// functions, so they are a counted named barrier (kBarCta, 96 threads) rather than __syncthreads(), whose contract asks
// every thread to reach the SAME call site.
enum : int { kBarO = 1, kBarZ = 2, kBarPow = 3, kBarCta = 4, kBarWin = 5, kBarEl = 6 };
enum : int { kFlagTone1 = 1, kFlagFirst = 2, kFlagLive = 4, kFlagExit = 8 };

// RING = samples per stream resident in shared memory (a multiple of 8).  The ring is what limits the number of
// resident CTAs per SM: 256 rows -> 45.0 KB -> 4 CTAs, 240 -> 5, 176 -> 6, 144 -> 7.
template <int RING, bool ELB = false>
struct __align__(16) BankSmem {
    static constexpr int kRingRows = RING;
    static constexpr bool kElb = ELB;
    uint32_t ring[RING + kMirrorRows][kSpc];
    double o[4][kSpc];            // WINDOW -> AFC: O1.r, O1.i, O2.r, O2.i
    double z[4][kSpc];            // AFC -> WINDOW: z1.r, z1.i, z2.r, z2.i
    double pw[10][kSpc];          // AFC -> WINDOW: q1, q2, qq1, qq2, zeta40
    double el[ELB ? 12 : 1][kSpc];  // AFC -> WINDOW (ELB): H0 (F1), H0 (F2), H5 (F1), H5 (F2), s0, s60
    int flags[kSpc];              // WINDOW -> AFC: kFlag*
    int w0[kSpc];                 // WINDOW -> STAGE: row-relative sample index of window slot 0 of the current symbol
    int fill[kSpc];               // STAGE -> WINDOW: samples [.., fill) of the stream's row are in the ring
    int live[kSpc];               // WINDOW -> STAGE: stream still has symbols to demodulate in this launch
    int exit_flag;                // WINDOW -> STAGE
};

template <int ID, int N>
__device__ __forceinline__ void bar_sync() { asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(N) : "memory"); }
template <int ID, int N>
__device__ __forceinline__ void bar_arrive() { asm volatile("bar.arrive %0, %1;" ::"n"(ID), "n"(N) : "memory"); }
__device__ __forceinline__ int ld_vol(const int* p) { return *reinterpret_cast<const volatile int*>(p); }
__device__ __forceinline__ void st_vol(int* p, int v) { *reinterpret_cast<volatile int*>(p) = v; }

// Ring word -> doubles.  I always converts with one I2F.F64.S16 (XU pipe: 8 cycles per warp instruction, one issue
// slot).  Q depends on the storage format, chosen by QX for the whole kernel:
//   QX = 0  Q stored offset-binary; 2^52 bias trick (two integer-pipe instructions + one DADD on the FP64 pipe)
//   QX = 1  Q stored offset-binary; odd window slots through XU (after undoing the offset), even slots bias trick
//   QX = 2  Q stored raw; one I2F.F64.S16 on the upper half (every conversion on the XU pipe)
// The kernel is bound by issue slots, the FP64 pipe and the XU pipe at nearly the same level; QX balances them (the
// launcher's default is QX = 0, 1 % ahead of the others on the full bank; no difference on the ELB path).
template <int QX>
__device__ __forceinline__ void unpack_ring(uint32_t w, int k, double& I, double& Q) {
    I = (double)(int16_t)(w & 0xFFFFu);
    if (QX == 2)
        Q = (double)(int16_t)(w >> 16);
    else if (QX == 1 && (k & 1))
        Q = (double)(int16_t)((w >> 16) ^ 0x8000u);
    else
        Q = __hiloint2double(0x43300000, (int)(w >> 16)) - 4503599627403264.0;  // 2^52 + 2^15
}
template <int QX>
__device__ __forceinline__ uint32_t ring_word(uint32_t raw) { return QX == 2 ? raw : raw ^ kQBias; }

__device__ __forceinline__ void ldg256(const uint32_t* p, uint4& a, uint4& b) {  // read-once, 32-byte aligned
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
                 : "l"(p));
}
__device__ __forceinline__ uint4 ldg128(const uint32_t* p) {
    uint4 a;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w)
                 : "l"(p));
    return a;
}

// 8 consecutive samples starting at rel (multiple of 8).  A ring is a multiple of 64 samples long, so a chunk never
// straddles its end; linear rows are 16-byte aligned and a multiple of 4 samples long, so only their last chunk can
// be partial.
__device__ __forceinline__ void chunk_load(const RowView& v, int rel, bool wide, uint4& a, uint4& b) {
    const uint32_t* p = v.row + v.phys(rel);
    if (rel + kChunk <= v.rel_end) {
        if (wide) {
            ldg256(p, a, b);
        } else {
            a = ldg128(p);
            b = ldg128(p + 4);
        }
    } else {
        a = (rel + 4 <= v.rel_end) ? ldg128(p) : make_uint4(0u, 0u, 0u, 0u);
        b = make_uint4(0u, 0u, 0u, 0u);
    }
}
// ring row of row-relative sample index idx (idx >= -RING: the first window of a stream starts kWinLead samples early)
template <int RING>
__device__ __forceinline__ int ring_row(int idx) {
    if ((RING & (RING - 1)) == 0) return idx & (RING - 1);
    return (int)((unsigned)(idx + RING) % (unsigned)RING);
}
template <int QX, class SM>
__device__ __forceinline__ void chunk_store(SM& sm, int s, int idx, uint4 a, uint4 b) {
    constexpr int kRingRows = SM::kRingRows;
    const int r = ring_row<kRingRows>(idx);
    const uint32_t w[8] = {ring_word<QX>(a.x), ring_word<QX>(a.y), ring_word<QX>(a.z), ring_word<QX>(a.w),
                           ring_word<QX>(b.x), ring_word<QX>(b.y), ring_word<QX>(b.z), ring_word<QX>(b.w)};
#pragma unroll
    for (int j = 0; j < 8; ++j) sm.ring[r + j][s] = w[j];
    if (r < kMirrorRows) {
#pragma unroll
        for (int j = 0; j < 8; ++j) sm.ring[kRingRows + r + j][s] = w[j];
    }
}

// early-gate correction for the first symbol of a call (rare: kept out of line)
template <int QX>
__device__ __noinline__ cplx first_fix_cold(const uint32_t* win, double f, cplx z) {
    return first_symbol_fix_w([&](int kk) { return ring_word<QX>(win[kk * kSpc]); }, f, z);
}
// call scheduling out of line; everything it touches by reference lives in local memory
__device__ __noinline__ bool schedule_cold(DemodState& st, int mode, long long avail, bool final_flag) {
    double pos = st.pos;
    const bool live = demod_schedule(st, pos, mode, avail, final_flag);
    st.pos = pos;
    return live;
}

// ---------------------------------------------------------------------------------------------------------------
// Per-stream bookkeeping of the window role that owns the timing loop and the call schedule.
// The DemodState record itself stays OUTSIDE this struct: its address escapes to the out-of-line scheduler, and an
// object that shares a struct with an address-taken member lands in local memory as a whole.
struct WindowCtl {
    long long avail, n_sym0, origin0, base_abs;
    double* soft_row;
    int soft_wrap, soft_idx, n_new;
    double timing_freq, pos, call_len_d, f;
    int sym_in_call, origin_rel, w0;
    bool live;

    __device__ __forceinline__ void init(DemodState& st, const StreamBuffers& sb, const SoftBuffers& so, int stream,
                                         bool valid, int mode, int final_flag) {
        avail = sb.avail[stream];
        soft_row = so.soft + (long long)stream * so.stride;
        soft_wrap = so.ring ? (int)so.stride : 0x7fffffff;
        soft_idx = (int)soft_pos(so, st.n_sym);  // row position of the next soft symbol
        n_new = 0;                               // symbols produced by this launch
        n_sym0 = st.n_sym; origin0 = st.origin;
        base_abs = make_row_view(sb, stream, st.origin).base_abs;
        timing_freq = st.timing_freq;
        live = valid && schedule_cold(st, mode, avail, final_flag != 0);
        pos = st.pos;
        sym_in_call = st.sym_in_call;
        call_len_d = (double)st.call_len;
        origin_rel = (int)(st.origin - base_abs);
        w0 = 0; f = 0.0;
        if (live) locate();
    }
    __device__ __forceinline__ void locate() {
        const int b = __double2int_rz(pos);  // pos >= 0: truncation == floor (:125)
        f = pos - (double)b;
        w0 = origin_rel + b - kWinLead;
    }
    __device__ __forceinline__ void put_soft(double v) {
        soft_row[soft_idx] = v;  // :268
        if (++soft_idx == soft_wrap) soft_idx = 0;
        ++n_new;
    }
    // after the timing loop has advanced pos: next symbol of this stream, maybe through a call boundary
    __device__ __forceinline__ void advance(DemodState& st, int mode, int final_flag) {
        sym_in_call = 1;  // any non-zero value: the open call has produced symbols
        if (!((pos + 40.0) + 10.0 < call_len_d)) {  // :221 fails: close the call, maybe open the next
            st.n_sym = n_sym0 + n_new;
            st.sym_in_call = sym_in_call;
            st.pos = pos;
            live = schedule_cold(st, mode, avail, final_flag != 0);
            pos = st.pos;
            sym_in_call = st.sym_in_call;
            call_len_d = (double)st.call_len;
            origin_rel = (int)(st.origin - base_abs);
        }
        if (live) locate();
    }
    // persist: this role writes its fields of the record, the AFC warp patches its own afterwards
    __device__ __forceinline__ void persist(DemodState& st, const SoftBuffers& so, DemodState* dstate, int stream,
                                            unsigned long long* counters) {
        st.n_sym = n_sym0 + n_new;
        so.n_sym[stream] = st.n_sym;
        DemodState* d = dstate + stream;
        d->pos = pos; d->timing_freq = timing_freq; d->origin = st.origin; d->call_len = st.call_len;
        d->n_sym = st.n_sym; d->sym_in_call = sym_in_call; d->flags = st.flags;
        unsigned long long dsym = (unsigned long long)n_new;
        unsigned long long dsmp = (unsigned long long)(st.origin - origin0);
        if (st.flags & kFlagDone) dsmp = (unsigned long long)(avail - origin0);
        if (dsym) atomicAdd(&counters[kCtrSymbols], dsym);
        if (dsmp) atomicAdd(&counters[kCtrSamples], dsmp);
    }
};

template <class SM>
__device__ __forceinline__ void wait_window(SM& sm, int s, bool live, int w0) {
    // normally true at once: the staging warp runs 2-4 symbols ahead
    while (!__all_sync(kFull, !live || ld_vol(&sm.fill[s]) >= w0 + kWin)) {}
    __threadfence_block();
}
template <class SM>
__device__ __forceinline__ void load_lo(const SM& sm, int s, BankLo& lo) {
    lo.z1 = {sm.z[0][s], sm.z[1][s]};
    lo.z2 = {sm.z[2][s], sm.z[3][s]};
    lo.inc1 = 0.0; lo.inc2 = 0.0;
}
template <class SM>
__device__ __forceinline__ void load_pow(const SM& sm, int s, BankPow& pw) {
    pw.q1 = {sm.pw[0][s], sm.pw[1][s]};
    pw.q2 = {sm.pw[2][s], sm.pw[3][s]};
    pw.qq1 = {sm.pw[4][s], sm.pw[5][s]};
    pw.qq2 = {sm.pw[6][s], sm.pw[7][s]};
    pw.zeta40 = {sm.pw[8][s], sm.pw[9][s]};
}

// ---------------------------------------------------------------------------------------------------------------
// AFC role.  NW = threads taking part in the z / power hand-offs (64: one window warp, 96: two).
// ELB: this warp also evaluates the early / late block sums H0, H5 of both tones (bank_el_blocks) while the window warp
// works on the on-time blocks; it is woken by kBarWin when the window of the symbol has been published.
template <int NW, int QX, class SM>
__device__ __forceinline__ void role_afc(SM& sm, int s, int stream, bool valid, DemodState* dstate, double afc_alpha) {
    const DemodState* d0 = dstate + stream;
    BankAfc afc = {d0->freq_offset, d0->ph1, d0->ph2, d0->p1, d0->p2};
    BankLo lo;
    BankPow pw;
    {
        double d;
        const cplx zeta = bank_zeta_general(afc.freq_offset, d);  // a -o offset may exceed the fast range
        bank_lo_from_zeta(zeta, d, lo, g_fm);
        bank_pow_from_zeta(zeta, pw, g_bk);
    }
    bar_sync<kBarCta, 96>();  // (1)
    auto publish_z = [&]() {
        sm.z[0][s] = lo.z1.r; sm.z[1][s] = lo.z1.i; sm.z[2][s] = lo.z2.r; sm.z[3][s] = lo.z2.i;
        bar_arrive<kBarZ, NW>();
    };
    auto publish_pow = [&]() {
        sm.pw[0][s] = pw.q1.r; sm.pw[1][s] = pw.q1.i; sm.pw[2][s] = pw.q2.r; sm.pw[3][s] = pw.q2.i;
        sm.pw[4][s] = pw.qq1.r; sm.pw[5][s] = pw.qq1.i; sm.pw[6][s] = pw.qq2.r; sm.pw[7][s] = pw.qq2.i;
        sm.pw[8][s] = pw.zeta40.r; sm.pw[9][s] = pw.zeta40.i;
        bar_arrive<kBarPow, NW>();
    };
    publish_z();
    publish_pow();
    for (;;) {
        if (SM::kElb) {
            bar_sync<kBarWin, 64>();  // the window of this symbol is published (or the launch is over)
            if (ld_vol(&sm.exit_flag)) break;
            const int w0 = ld_vol(&sm.w0[s]);
            wait_window(sm, s, ld_vol(&sm.live[s]) != 0, w0);
            const uint32_t* const win = &sm.ring[ring_row<SM::kRingRows>(w0)][s];
            auto slot = [&](int k, double& I, double& Q) { unpack_ring<QX>(win[k * kSpc], k, I, Q); };
            BankElBlocks e;
            bank_el_blocks(slot, lo.z1, lo.z2, e);
            sm.el[0][s] = e.H0a.r; sm.el[1][s] = e.H0a.i; sm.el[2][s] = e.H0b.r; sm.el[3][s] = e.H0b.i;
            sm.el[4][s] = e.H5a.r; sm.el[5][s] = e.H5a.i; sm.el[6][s] = e.H5b.r; sm.el[7][s] = e.H5b.i;
            sm.el[8][s] = e.s0.r; sm.el[9][s] = e.s0.i; sm.el[10][s] = e.s60.r; sm.el[11][s] = e.s60.i;
            bar_arrive<kBarEl, 64>();
        }
        bar_sync<kBarO, 64>();
        const int fl = sm.flags[s];
        if (fl & kFlagExit) break;  // warp-uniform: the window warp sets it on every lane
        const cplx O1 = {sm.o[0][s], sm.o[1][s]}, O2 = {sm.o[2][s], sm.o[3][s]};
        // no AFC update on the first symbol of a call (:289): same LO steps next symbol
        const bool update = (fl & kFlagLive) && !(fl & kFlagFirst);
        cplx zeta = {1.0, 0.0};
        if (fl & kFlagLive) {
            bank_afc(afc, O1, O2, (fl & kFlagTone1) != 0, pw.zeta40, lo.inc1, lo.inc2, (fl & kFlagFirst) != 0, afc_alpha,
                     g_fm);
            if (update) {
                double d;
                zeta = bank_zeta_fast(afc.freq_offset, d, g_fm);  // |offset| <= 2 kHz after the clamp (:303)
                bank_lo_from_zeta(zeta, d, lo, g_fm);
            }
        }
        publish_z();  // z goes out first: the next symbol's Horner needs nothing else
        if (update) bank_pow_from_zeta(zeta, pw, g_bk);
        publish_pow();
    }
    bar_sync<kBarCta, 96>();  // (2) the window warp has written the records
    if (valid) {
        DemodState* d = dstate + stream;
        d->freq_offset = afc.freq_offset; d->ph1 = afc.ph1; d->ph2 = afc.ph2; d->p1 = afc.p1; d->p2 = afc.p2;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// STAGE role.  NB chunks (32-byte sectors) per batch and lane, two batches in flight.
template <int QX, int NB, class SM>
__device__ __forceinline__ void role_stage(SM& sm, int s, int stream, const StreamBuffers& sb, const DemodState* dstate) {
    constexpr int kBatch = kChunk * NB;
    constexpr int kRingRows = SM::kRingRows;
    // a batch in flight is urgent when the window could need it before the next round (about half a symbol later)
    constexpr int kUrgent = kRingRows >= 240 ? kWin + 2 * kBatch : kWin + 40 + kChunk;
    const RowView view = make_row_view(sb, stream, dstate[stream].origin);  // same base as the window warp
    const bool wide = ((reinterpret_cast<uintptr_t>(sb.iq) | (uintptr_t)(sb.stride * 4)) & 31u) == 0;
    sm.fill[s] = -(1 << 30);
    bar_sync<kBarCta, 96>();  // (1)
    int req;  // samples [.., req) of this lane's row have been requested (multiple of 8)
    {
        const int w0 = sm.w0[s];
        req = (w0 < 0 ? 0 : w0) & ~(kChunk - 1);
    }
    int pub = req;  // samples [.., pub) are in the ring and published in sm.fill (stored in request order: one watermark)
    uint4 buf[2 * NB];
    int idx = -1;   // row index of the batch in flight (-1: none)
    // No blocking primitive here: the warp paces itself with __nanosleep.  (A named-barrier wake-up was tried: a blocked
    // warp costs nothing, but one ordering slip between the watermark and the barrier deadlocks the CTA, and a hung
    // kernel on a shared box costs more than the ~25 polling instructions per symbol this loop spends.  The first version
    // polled every 200 ns with a heavy loop body and spent ~670 instructions per symbol, on a sub-partition it shares
    // with another CTA's window warp.)  Steady state per round: store the batch requested before the last sleep (its
    // loads have had a sleep and a round to land), request the next one, sleep.  A stream that is close to starving (start of a
    // launch, end of a row, catching up after a stall) is served without sleeping.
    for (;;) {
        const int w0 = ld_vol(&sm.w0[s]);
        const bool lv = ld_vol(&sm.live[s]) != 0;
        if (idx >= 0) {
#pragma unroll
            for (int c = 0; c < NB; ++c) chunk_store<QX>(sm, s, idx + kChunk * c, buf[2 * c], buf[2 * c + 1]);
            __threadfence_block();
            pub = idx + kBatch;
            st_vol(&sm.fill[s], pub);
            idx = -1;
        }
        if (lv && req + kBatch <= w0 + kRingRows - kChunk && req < view.rel_end) {
#pragma unroll
            for (int c = 0; c < NB; ++c) chunk_load(view, req + kChunk * c, wide, buf[2 * c], buf[2 * c + 1]);
            idx = req;
            req += kBatch;
        }
        const bool urgent = idx >= 0 && pub < w0 + kUrgent;  // the window may need this batch before the next round
        if (__any_sync(kFull, urgent)) continue;
        if (ld_vol(&sm.exit_flag)) break;
        __nanosleep(g_stage_sleep_ns);
    }
    bar_sync<kBarCta, 96>();  // (2)
}

}  // namespace

// =================================================================================================================
// Three-warp kernel: WINDOW, AFC, STAGE (96 threads)
// RING / MINB: ring rows per stream and the resident CTAs per SM they allow.  The register cap is written out
// (65,536 / (96 MINB) rounded down to the allocation unit of 8): __launch_bounds__' own choice is more conservative.
template <int QX, int RING, int MINB, bool ELB>
__global__ void __maxnreg__(MINB == 4 ? 168 : MINB == 5 ? 136 : MINB == 6 ? 112 : 96)
demod_bank_kernel(StreamBuffers sb, SoftBuffers so, DemodState* __restrict__ dstate, int n_streams, int mode,
                  int final_flag, double afc_alpha, unsigned long long* __restrict__ counters) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using SM = BankSmem<RING, ELB>;
    SM& sm = *reinterpret_cast<SM*>(smem_raw);
    const int s = threadIdx.x & 31, role = threadIdx.x >> 5;
    const int stream_raw = blockIdx.x * kSpc + s;
    const bool valid = stream_raw < n_streams;
    const int stream = valid ? stream_raw : n_streams - 1;

    if (role == 1) { role_afc<64, QX>(sm, s, stream, valid, dstate, afc_alpha); return; }
    if (role == 2) { role_stage<QX, (RING >= 176 ? 5 : 3)>(sm, s, stream, sb, dstate); return; }

    DemodState st = dstate[stream];  // local memory: only the scheduler touches it
    WindowCtl c;
    c.init(st, sb, so, stream, valid, mode, final_flag);
    sm.w0[s] = c.w0;
    sm.live[s] = c.live ? 1 : 0;
    if (s == 0) sm.exit_flag = 0;
    bar_sync<kBarCta, 96>();  // (1) symbol 0 published
    bool any_live = __any_sync(kFull, c.live);
    while (any_live) {
        const bool first = c.sym_in_call == 0;
        if (ELB) bar_arrive<kBarWin, 64>();  // the AFC warp may start on the early / late blocks of this window
        bar_sync<kBarZ, 64>();
        BankLo lo;
        load_lo(sm, s, lo);
        wait_window(sm, s, c.live, c.w0);
        const uint32_t* const win = &sm.ring[ring_row<RING>(c.w0)][s];
        auto slot = [&](int k, double& I, double& Q) { unpack_ring<QX>(win[k * kSpc], k, I, Q); };
        cplx A[4], B[4], s10, s20, s40;
        bank_on_blocks(slot, lo.z1, lo.z2, A, B, s10, s20, s40);
        bar_sync<kBarPow, 64>();
        BankPow pw;
        load_pow(sm, s, pw);
        BankOnTime on;
        bank_on_time(slot, c.f, lo, pw, A, B, s10, s20, s40, on);
        const bool tone1 = on.eO1 > on.eO2;  // :272, :291
        sm.o[0][s] = on.O1.r; sm.o[1][s] = on.O1.i; sm.o[2][s] = on.O2.r; sm.o[3][s] = on.O2.i;
        sm.flags[s] = (tone1 ? kFlagTone1 : 0) | (first ? kFlagFirst : 0) | (c.live ? kFlagLive : 0);
        bar_arrive<kBarO, 64>();  // the AFC warp takes it from here
        if (c.live) c.put_soft(on.eO2 - on.eO1);
        cplx fixE = {0.0, 0.0};
        if (first && c.live) fixE = first_fix_cold<QX>(win, c.f, tone1 ? lo.z1 : lo.z2);  // :237, once per call
        double eE, eL;
        if (ELB) {
            bar_sync<kBarEl, 64>();  // block sums H0, H5 of both tones from the AFC warp
            const int t = tone1 ? 0 : 2;
            const cplx H0 = {sm.el[t][s], sm.el[t + 1][s]}, H5 = {sm.el[4 + t][s], sm.el[5 + t][s]};
            const cplx s0 = {sm.el[8][s], sm.el[9][s]}, s60 = {sm.el[10][s], sm.el[11][s]};
            bank_early_late_from_blocks(c.f, tone1, lo, pw, on, H0, H5, s0, s60, fixE, eE, eL);
        } else {
            bank_early_late(slot, c.f, tone1, lo, pw, on, fixE, eE, eL);
        }
        if (c.live) {
            bank_timing(eE, eL, c.timing_freq, c.pos, g_fm);
            c.advance(st, mode, final_flag);
            if (c.live) st_vol(&sm.w0[s], c.w0);
            else st_vol(&sm.live[s], 0);
        }
        any_live = __any_sync(kFull, c.live);
    }
    sm.flags[s] = kFlagExit;
    st_vol(&sm.exit_flag, 1);
    if (ELB) bar_arrive<kBarWin, 64>();  // the AFC warp waits for the next window
    else bar_arrive<kBarO, 64>();        // ... or for the next on-time correlations
    if (valid) c.persist(st, so, dstate, stream, counters);
    bar_sync<kBarCta, 96>();  // (2)
}

template <int QX, int RING, int MINB, bool ELB>
static cudaError_t launch_bank_t(const StreamBuffers& sb, const SoftBuffers& so, DemodState* dstate, int n_streams, int mode,
                                 int final_flag, double afc_alpha, unsigned long long* counters, cudaStream_t st) {
    const size_t smem = sizeof(BankSmem<RING, ELB>);
    cudaError_t e = cudaFuncSetAttribute(demod_bank_kernel<QX, RING, MINB, ELB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    // all of the SM's shared memory, or the resident CTAs the ring was sized for do not fit
    e = cudaFuncSetAttribute(demod_bank_kernel<QX, RING, MINB, ELB>, cudaFuncAttributePreferredSharedMemoryCarveout,
                             (int)cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    const int grid = (n_streams + kSpc - 1) / kSpc;
    demod_bank_kernel<QX, RING, MINB, ELB><<<grid, 96, smem, st>>>(sb, so, dstate, n_streams, mode, final_flag, afc_alpha, counters);
    return cudaGetLastError();
}
// Measured and dropped: more resident CTAs per SM through smaller rings and register caps
// (RING 240/176/144 with 5/6/7 CTAs): a 176-row ring alone costs +36 % (too little prefetch distance for HBM latency),
// a 112-register cap alone +23 %, and banks sized for 5-7 CTAs per SM ran slower in aggregate than two waves of four.
cudaError_t launch_demod_bank(const StreamBuffers& sb, const SoftBuffers& so, DemodState* dstate, int n_streams,
                              int mode, int final_flag, double afc_alpha, unsigned long long* counters,
                              cudaStream_t st) {
    // ELB (the AFC warp takes the early / late block sums) pays while SM sub-partitions are idle: up to one CTA per SM
    // (probe banks x 6 frames: 4,736 streams 14.4 against 15.9 ms; 9,472 streams 18.7 against 16.9 ms)
    int sms = 148, dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    static const int force = getenv("OPVD_BANK_ELB") ? atoi(getenv("OPVD_BANK_ELB")) : -1;  // development switch
    const bool elb = force >= 0 ? force != 0 : (n_streams + kSpc - 1) / kSpc <= sms;
    static const int sleep_ns = getenv("OPVD_BANK_SLEEP") ? atoi(getenv("OPVD_BANK_SLEEP")) : 0;
    if (sleep_ns > 0) {
        const unsigned v = (unsigned)sleep_ns;
        cudaMemcpyToSymbolAsync(g_stage_sleep_ns, &v, sizeof(v), 0, cudaMemcpyHostToDevice, st);
    }
    // Q conversion split (unpack_ring): every split gives the same values; on the full bank QX = 0 measured 19.84 ms,
    // QX = 1 20.03 ms, QX = 2 20.24 ms (profiles/gpu_r02_qx.log).  OPVD_BANK_QX: development switch.
    static const int qx = getenv("OPVD_BANK_QX") ? atoi(getenv("OPVD_BANK_QX")) : 0;
    if (elb) return launch_bank_t<0, 256, 4, true>(sb, so, dstate, n_streams, mode, final_flag, afc_alpha, counters, st);
    if (qx == 1) return launch_bank_t<1, 256, 4, false>(sb, so, dstate, n_streams, mode, final_flag, afc_alpha, counters, st);
    if (qx == 2) return launch_bank_t<2, 256, 4, false>(sb, so, dstate, n_streams, mode, final_flag, afc_alpha, counters, st);
    return launch_bank_t<0, 256, 4, false>(sb, so, dstate, n_streams, mode, final_flag, afc_alpha, counters, st);
}

}  // namespace opvd
