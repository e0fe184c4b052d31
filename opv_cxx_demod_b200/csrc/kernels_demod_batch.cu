// kernels_demod_batch.cu — A1/A3/A4 for sm_100a, BATCHED variant for large channel banks
// (demod_batch_core.cuh).  Used when there are enough streams to fill the machine (thousands):
// the goal is throughput per issued instruction, not per-symbol latency (kernels_demod_warp.cu).
//
// A CTA of 160 threads (5 specialised warps) owns 32 streams; lane = stream in every warp.  Every
// symbol has two phases separated by CTA barriers:
//   window phase  helper warps k = 0..3: tone k >> 1, window half k & 1.  30 samples -> three
//                 10-sample Horner block sums -> three partial gates to shared memory.  There is no
//                 intra-warp exchange and every instruction does 32 streams' worth of work.
//                 The post-sum interpolator is linear, so each helper applies it to its own partials.
//   loop phase    loop warp (warp 4): combines the halves (3 complex FMAs per tone), soft decision, early-late timing loop, AFC (branch-free atan2), LO steps for the
//                 next symbol, call schedule.  Helper warps 1-3 meanwhile move the next samples from
//                 HBM into the shared-memory ring.
// Separate warps keep the loop state out of the helpers' register budget (and vice versa).
// With one lane per stream in the loop phase the serial arithmetic of the recurrence is amortised
// over 32 streams (the warp-per-stream kernel spends 330 warp-instructions per stream and symbol,
// this kernel ~45), and with four threads per stream in the window phase a 16,384-stream bank puts
// 16 warps on every SM instead of 3.5.
//
// Sample ring: transposed, ring[row][stream] with row = sample index mod 256, so that lane s always
// reads bank s whatever its stream's window position is (per-stream rings laid out stream-major
// give 3-4-way bank conflicts on every load because the window offsets of the 32 streams are
// unrelated).  Rows 0..63 are mirrored behind row 255: a 61-row window never wraps.  Warps 1-3 fill
// it with 128-bit global loads (each thread 64 contiguous bytes of its stream per symbol, issued one
// full symbol before they are stored), so HBM is read exactly once, in whole 32-byte sectors.
#include <cuda_runtime.h>
#include <cstdint>

#include "demod_batch_core.cuh"
#include "demod_warp_core.cuh"  // first_symbol_fix_w
#include "opvd_kernels.cuh"

namespace opvd {

namespace {

constexpr int kSpc = 32;            // streams per CTA
constexpr int kHelperWarps = 4;     // window-phase roles per stream
constexpr int kThreads = 32 * (kHelperWarps + 1);  // + the loop warp
constexpr int kRingRows = 256;      // samples per stream resident in shared memory (power of two)
constexpr int kMirrorRows = 64;     // rows 0..63 repeated after row 255
constexpr int kRows = kRingRows + kMirrorRows;
constexpr int kStage = 16;          // samples per staging thread and symbol (4 x LDG.128)
constexpr int kStageAll = 3 * kStage;  // per stream and symbol

struct __align__(16) BatchSmem {
    uint32_t ring[kRows][kSpc];   // 40 KB
    double2 part[4][3][kSpc];     // [role][E,O,L][stream] partial gates, 6 KB
    double zq[2][4][kSpc];        // [tone][z.r, z.i, q.r, q.i][stream], 2 KB
    double frac[kSpc];            // interpolation fraction f = pos - floor(pos) of the current symbol
    int w0[kSpc];                 // row-relative sample index of window slot 0 of the current symbol
    int live[kSpc];               // stream has a symbol to demodulate
    int any_live;
};

__device__ __forceinline__ void stage_store(BatchSmem& sm, int s, int idx, const uint4 (&v)[4]) {
    // 16 consecutive samples of stream s starting at sample index idx (multiple of 16)
    const int row = idx & (kRingRows - 1);
    const uint32_t w[16] = {v[0].x, v[0].y, v[0].z, v[0].w, v[1].x, v[1].y, v[1].z, v[1].w,
                            v[2].x, v[2].y, v[2].z, v[2].w, v[3].x, v[3].y, v[3].z, v[3].w};
#pragma unroll
    for (int j = 0; j < 16; ++j) sm.ring[row + j][s] = w[j];
    if (row < kMirrorRows) {
#pragma unroll
        for (int j = 0; j < 16; ++j) sm.ring[kRingRows + row + j][s] = w[j];
    }
}

// 16 samples of a row starting at idx (multiple of 16); rows are 16-byte aligned and a multiple of 4
// samples long, so only the last chunk of a row can be partial
__device__ __forceinline__ void stage_load(const uint32_t* row, int idx, int stride, uint4 (&v)[4]) {
    const uint4* p = reinterpret_cast<const uint4*>(row + idx);
    if (idx + kStage <= stride) {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = __ldg(p + j);
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = (idx + 4 * j + 4 <= stride) ? __ldg(p + j) : make_uint4(0u, 0u, 0u, 0u);
    }
}

// early-gate correction for the first symbol of a call, both tones (rare: kept out of line)
__device__ __noinline__ cplx first_fix_cold(const uint32_t* win, double f, cplx z) {
    return first_symbol_fix_w([&](int kk) { return win[kk * kSpc]; }, f, z);
}
// call scheduling out of line.  Everything it touches by reference lives in local memory, so the
// hot loop hands it copies: the position comes back through st.pos.
__device__ __noinline__ bool schedule_cold(DemodState& st, int mode, long long avail, bool final_flag) {
    double pos = st.pos;
    const bool live = demod_schedule(st, pos, mode, avail, final_flag);
    st.pos = pos;
    return live;
}

__device__ __forceinline__ void publish_lo(BatchSmem& sm, int s, const BatchRegs& r) {
    sm.zq[0][0][s] = r.t1.z.r; sm.zq[0][1][s] = r.t1.z.i; sm.zq[0][2][s] = r.t1.q.r; sm.zq[0][3][s] = r.t1.q.i;
    sm.zq[1][0][s] = r.t2.z.r; sm.zq[1][1][s] = r.t2.z.i; sm.zq[1][2][s] = r.t2.q.r; sm.zq[1][3][s] = r.t2.q.i;
}

// ------------------------------------------------------------------------------------------------
// loop warp: one lane per stream
__device__ __forceinline__ void loop_warp(BatchSmem& sm, const StreamBuffers& sb, const SoftBuffers& so,
                                          DemodState* __restrict__ dstate, int stream, bool valid, int s, int mode,
                                          int final_flag, double afc_alpha, unsigned long long* __restrict__ counters) {
    const long long row0 = sb.row_base;
    DemodState st = dstate[stream];
    const long long avail = sb.avail[stream];
    double* const soft_row = so.soft + (long long)stream * so.stride - so.base;
    BatchRegs r;
    r.freq_offset = st.freq_offset; r.ph1 = st.ph1; r.ph2 = st.ph2; r.pos = st.pos; r.timing_freq = st.timing_freq;
    r.p1 = st.p1; r.p2 = st.p2;
    batch_lo(r.freq_offset, r.t1, r.t2);  // general version: a -o offset may exceed the fast range
    const long long n_sym0 = st.n_sym, origin0 = st.origin;
    bool live = valid && schedule_cold(st, mode, avail, final_flag != 0);
    r.pos = st.pos;
    double call_len_d = (double)st.call_len, f = 0.0;
    int origin_rel = (int)(st.origin - row0);
    int w0 = 0;
    // st lives in local memory (its address goes to the out-of-line scheduler): keep the per-symbol
    // counters in registers and write them back only around that call
    int sym_in_call = st.sym_in_call;
    double* soft_ptr = soft_row + st.n_sym;
    if (live) {
        const int b = __double2int_rz(r.pos);  // pos >= 0: truncation == floor (:125)
        f = r.pos - (double)b;
        w0 = origin_rel + b - kWinLead;
    }
    sm.w0[s] = w0;
    sm.frac[s] = f;
    sm.live[s] = live ? 1 : 0;
    publish_lo(sm, s, r);
    {
        const int any = __any_sync(0xffffffffu, live);
        if (s == 0) sm.any_live = any;
    }
    __syncthreads();  // state of symbol 0 published
    __syncthreads();  // ring primed by the helpers

    while (sm.any_live) {
        const bool first = sym_in_call == 0;
        __syncthreads();  // partial gates ready
        if (live) {
            // one tone at a time (the barrier keeps the second tone's loads from being hoisted, which
            // would double the live registers)
            const uint32_t* win = &sm.ring[w0 & (kRingRows - 1)][s];  // window n is still in the ring
            ToneGates g1, g2;
            {
                HalfGates a, b;
                double2 v;
                v = sm.part[0][0][s]; a.E = {v.x, v.y}; v = sm.part[0][1][s]; a.O = {v.x, v.y}; v = sm.part[0][2][s]; a.L = {v.x, v.y};
                v = sm.part[1][0][s]; b.E = {v.x, v.y}; v = sm.part[1][1][s]; b.O = {v.x, v.y}; v = sm.part[1][2][s]; b.L = {v.x, v.y};
                cplx fix = {0.0, 0.0};
                if (first) fix = first_fix_cold(win, f, r.t1.z);  // early-gate clamp (:237), once per call
                g1 = batch_finish_tone(a, b, r.t1, fix);
            }
            asm volatile("" ::: "memory");
            {
                HalfGates a, b;
                double2 v;
                v = sm.part[2][0][s]; a.E = {v.x, v.y}; v = sm.part[2][1][s]; a.O = {v.x, v.y}; v = sm.part[2][2][s]; a.L = {v.x, v.y};
                v = sm.part[3][0][s]; b.E = {v.x, v.y}; v = sm.part[3][1][s]; b.O = {v.x, v.y}; v = sm.part[3][2][s]; b.L = {v.x, v.y};
                cplx fix = {0.0, 0.0};
                if (first) fix = first_fix_cold(win, f, r.t2.z);
                g2 = batch_finish_tone(a, b, r.t2, fix);
            }
            const double soft = batch_symbol_serial(r, g1, g2, first, afc_alpha, g_fm);
            *soft_ptr++ = soft;
            sym_in_call = 1;  // any non-zero value: the open call has produced symbols
            // ---- next symbol of this stream
            if (!((r.pos + 40.0) + 10.0 < call_len_d)) {  // :221 fails: close the call, maybe open the next
                st.n_sym = (long long)(soft_ptr - soft_row);
                st.sym_in_call = sym_in_call;
                st.pos = r.pos;
                live = schedule_cold(st, mode, avail, final_flag != 0);
                r.pos = st.pos;
                sym_in_call = st.sym_in_call;
                call_len_d = (double)st.call_len;
                origin_rel = (int)(st.origin - row0);
            }
            if (live) {
                const int b2 = __double2int_rz(r.pos);
                f = r.pos - (double)b2;
                w0 = origin_rel + b2 - kWinLead;
                sm.w0[s] = w0;
                sm.frac[s] = f;
                publish_lo(sm, s, r);
            } else {
                sm.live[s] = 0;
            }
        }
        const int any = __any_sync(0xffffffffu, live);
        if (s == 0) sm.any_live = any;
        __syncthreads();  // state of the next symbol published, ring advanced
    }

    // ---- persist the stream's state
    if (valid) {
        st.n_sym = (long long)(soft_ptr - soft_row);
        st.sym_in_call = sym_in_call;
        st.freq_offset = r.freq_offset; st.ph1 = r.ph1; st.ph2 = r.ph2; st.pos = r.pos; st.timing_freq = r.timing_freq;
        st.p1 = r.p1; st.p2 = r.p2;
        dstate[stream] = st;
        unsigned long long dsym = (unsigned long long)(st.n_sym - n_sym0);
        unsigned long long dsmp = (unsigned long long)(st.origin - origin0);
        if (st.flags & kFlagDone) dsmp = (unsigned long long)(avail - origin0);
        if (dsym) atomicAdd(&counters[kCtrSymbols], dsym);
        if (dsmp) atomicAdd(&counters[kCtrSamples], dsmp);
    }
}

// ------------------------------------------------------------------------------------------------
// helper warps: window phase (all four) and ring staging (warps 1-3)
__device__ __forceinline__ void helper_warp(BatchSmem& sm, const uint32_t* __restrict__ row, int stride_i, int s, int k) {
    __syncthreads();  // state of symbol 0 published
    // ---- prime the ring: everything up to w0 + 208.. of each live stream
    int fill;  // samples [.., fill) of this thread's stream have been requested (multiple of 16)
    {
        const int w0 = sm.w0[s];
        fill = (w0 < 0 ? 0 : w0) & ~(kStage - 1);
        if (k >= 1 && sm.live[s]) {
            while (fill + kStageAll <= w0 + kRingRows) {
                const int idx = fill + kStage * (k - 1);
                if (idx < stride_i) {
                    uint4 v[4];
                    stage_load(row, idx, stride_i, v);
                    stage_store(sm, s, idx, v);
                }
                fill += kStageAll;
            }
        }
    }
    __syncthreads();  // ring primed

    uint4 pend[4];  // 16 samples requested during the previous symbol, stored during this one
    int pend_idx = -1;
    const int tone = k >> 1, half = k & 1;
    while (sm.any_live) {
        const int lv = sm.live[s];
        const int w0 = sm.w0[s];
        // ---- window phase
        if (lv) {
            const uint32_t* src = &sm.ring[(w0 & (kRingRows - 1)) + 30 * half][s];
            double I[31], Q[31];  // slots 30h .. 30h+29, and slot 60 for the late gate's edge term (h = 1)
#pragma unroll
            for (int j = 0; j < 30; ++j) unpack_iq_mixed(src[j * kSpc], I[j], Q[j]);
            I[30] = 0.0; Q[30] = 0.0;
            if (half) unpack_iq(src[30 * kSpc], I[30], Q[30]);
            const cplx z = {sm.zq[tone][0][s], sm.zq[tone][1][s]}, q = {sm.zq[tone][2][s], sm.zq[tone][3][s]};
            const HalfGates g = batch_half_gates(I, Q, z, q, sm.frac[s], half);
            sm.part[k][0][s] = make_double2(g.E.r, g.E.i);
            sm.part[k][1][s] = make_double2(g.O.r, g.O.i);
            sm.part[k][2][s] = make_double2(g.L.r, g.L.i);
        }
        __syncthreads();  // partial gates ready
        // ---- staging (warps 1-3): store the 16 samples requested one symbol ago (their rows hold samples
        // older than any live window), then request the next ones; the loads have a whole symbol to land
        if (k >= 1) {
            if (pend_idx >= 0) stage_store(sm, s, pend_idx, pend);
            pend_idx = -1;
            if (lv && fill + kStageAll <= w0 + kRingRows) {
                const int idx = fill + kStage * (k - 1);
                if (idx < stride_i) {
                    stage_load(row, idx, stride_i, pend);
                    pend_idx = idx;
                }
                fill += kStageAll;
            }
        }
        __syncthreads();  // state of the next symbol published, ring advanced
    }
}

}  // namespace

__global__ void __launch_bounds__(kThreads, 4)
demod_batch_kernel(StreamBuffers sb, SoftBuffers so, DemodState* __restrict__ dstate, int n_streams, int mode,
                   int final_flag, double afc_alpha, unsigned long long* __restrict__ counters) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BatchSmem& sm = *reinterpret_cast<BatchSmem*>(smem_raw);
    const int s = threadIdx.x & 31, k = threadIdx.x >> 5;
    const int stream_raw = blockIdx.x * kSpc + s;
    const bool valid = stream_raw < n_streams;
    const int stream = valid ? stream_raw : n_streams - 1;
    if (k == kHelperWarps) {
        loop_warp(sm, sb, so, dstate, stream, valid, s, mode, final_flag, afc_alpha, counters);
    } else {
        helper_warp(sm, sb.iq + (long long)stream * sb.stride, (int)sb.stride, s, k);
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
