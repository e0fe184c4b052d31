// kernels_decode.cu — A6/A7/A8: frame decoder, one warp per frame.
//   scale + 3-bit quantise           FrameDecoder::decode  /root/reference/src/opv-demod.cpp:856-866
//   67x32 bit-reversed deinterleave  deinterleave_addr     :792-795   (fused into the Viterbi input load)
//   K=7 r=1/2 Viterbi                ViterbiDecoder::decode :802-846
//   pack + LFSR derandomise          :878-895             (fused into the traceback epilogue)
//
// Viterbi layout: 64 states, 2 per lane.  Lane j holds the path metrics of states j and j+32
// (the two predecessors of states 2j and 2j+1), packed as 2 x int16 in one register, so the
// add-compare-select butterfly is lane-local: the two packed branch metrics of the step come from a
// 64-entry table (indexed by the step's two 3-bit symbols, XORed per lane with the lane's expected
// coded bits), two 32-bit adds form all four candidates, one packed signed 16-bit minimum does both
// compare-selects, and the two decisions are the sign bits of one guarded packed subtraction, which
// gives the reference's tie rule (a <= b keeps predecessor p0, :829).  The new metrics are
// redistributed with two warp shuffles + two byte-permutes.  13 instructions per trellis step (the
// kernel is issue-bound on the integer pipe).  Metrics are bounded by 2144*7 = 15,008, so int16 holds
// them; the reference's INT_MAX "unreachable" marker (:805,:826) is replaced by 16,384, which can
// never win against a reachable path and never reaches the final traceback.
// Decisions (64 bits/step) live in shared memory; traceback, byte packing and the derandomiser XOR
// run on lane 0 and the 134 bytes leave with one coalesced store.
#include <cuda_runtime.h>
#include <cstdint>

#include "opvd_kernels.cuh"

namespace opvd {

__constant__ uint8_t c_lfsr[kFrameBytes];

static void make_lfsr(uint8_t* out) {  // :887-893
    uint8_t lfsr = 0xFF;
    for (int i = 0; i < kFrameBytes; ++i) {
        uint8_t r = 0;
        for (int b = 7; b >= 0; --b) {
            r |= (uint8_t)(((lfsr >> 7) & 1) << b);
            lfsr = (uint8_t)((lfsr << 1) | (((lfsr >> 7) ^ (lfsr >> 6) ^ (lfsr >> 4) ^ (lfsr >> 2)) & 1));
        }
        out[i] = r;
    }
}

void upload_constants() {
    uint8_t t[kFrameBytes];
    make_lfsr(t);
    cudaMemcpyToSymbol(c_lfsr, t, sizeof(t));
}

constexpr int kDecWarps = 2;  // 64-thread CTAs (22 KB, 4,096 registers): one fits beside four resident channel-bank CTAs, so the
                              // decoder of time tile t runs while tile t+1 is demodulated (opvd_api.cu, back stream)
constexpr int kDecWords = kFrameBits / 16;                 // 67 decision words per lane
constexpr int kDecBytesPerWarp = kDecWords * 32 * 4;       // 8576 B, aliased as 1072 doubles for the scale sum
constexpr int kQBytesPerWarp = kEncodedBits;               // deinterleaved 3-bit symbols, one byte each
constexpr int kOutBytesPerWarp = 16;                      // (the 134 output bytes alias the symbol area, dead by then)
constexpr int kDecSmemPerWarp = kDecBytesPerWarp + kQBytesPerWarp + kOutBytesPerWarp;  // 10,864 B
constexpr int kBmTableBytes = 64 * 8;                      // per CTA: packed branch metrics by (u1, u2)
constexpr int kUnreachable = 16384;

// Branch-metric table, shared by the CTA: entry (u1 | u2 << 3) holds, for a butterfly whose predecessor
// p0 expects coded bits whose costs are u1 and u2 (:823-824 with e1/e2 folded in by the caller's XOR),
//   x = A0 | (14 - A0) << 16,  A0 = u1 + u2         p0 = state j,    input 0 | input 1 (both coded bits flip)
//   y = B0 | (14 - B0) << 16,  B0 = u1 + (u2 ^ 7)   p1 = state j+32 (G2 taps register bit 5, G1 does not)
__device__ __forceinline__ void build_bm_table(uint2* tbl) {
    for (int e = threadIdx.x; e < 64; e += blockDim.x) {
        const uint32_t u1 = e & 7, u2 = e >> 3;
        const uint32_t A0 = u1 + u2, B0 = u1 + (u2 ^ 7u);
        tbl[e] = make_uint2(A0 * 0xFFFF0001u + 0x000E0000u, B0 * 0xFFFF0001u + 0x000E0000u);
    }
    __syncthreads();
}

// inverse of deinterleave_addr: deint[i] = q[deinterleave_addr(i)]  =>  q index j feeds deint index inv(j)
__device__ __forceinline__ int deinterleave_inv(int j) {
    const int pos = (j & ~7) + 7 - (j & 7);
    return (pos % 67) * 32 + pos / 67;
}

// soft: start of the stream's soft row, p0: row position of the first payload symbol, wrap: ring length of the row
// (a payload of 2144 symbols may wrap once; 2^62 for a linear row)
__device__ __forceinline__ void decode_one(const double* __restrict__ soft_row, long long p0, long long wrap, unsigned char* wsm,
                                           const uint2* bm_tbl, uint8_t* out_frame, int32_t* out_metric,
                                           unsigned long long* counters) {
    auto soft_at = [&](int j) {
        long long q = p0 + j;
        if (q >= wrap) q -= wrap;
        return soft_row[q];
    };
    const int lane = threadIdx.x & 31;
    double* dsum = reinterpret_cast<double*>(wsm);
    uint32_t* dec = reinterpret_cast<uint32_t*>(wsm);
    uint8_t* q = wsm + kDecBytesPerWarp;
    uint8_t* outb = q;  // written after the forward pass has consumed the symbols

    // ---- scale = mean |soft| with the reference's sequential summation order (:856-858)
    double scale = 0.0;
    for (int half = 0; half < 2; ++half) {
        for (int i = lane; i < kEncodedBits / 2; i += 32) dsum[i] = fabs(soft_at(half * (kEncodedBits / 2) + i));
        __syncwarp();
#pragma unroll 8
        for (int i = 0; i < kEncodedBits / 2; ++i) scale += dsum[i];  // lane-uniform broadcast reads
        __syncwarp();
    }
    scale /= (double)kEncodedBits;
    if (scale < 1e-10) {  // :859 frame dropped
        if (lane == 0) {
            *out_metric = -1;
            atomicAdd(&counters[kCtrFramesDropped], 1ull);
        }
        for (int i = lane; i < kFrameBytes; i += 32) out_frame[i] = 0;
        return;
    }
    // ---- quantise (:863-866) and scatter to deinterleaved order (:869-871)
    for (int j = lane; j < kEncodedBits; j += 32) {
        const double n = __dadd_rn(__dmul_rn(__ddiv_rn(-soft_at(j), scale), 3.5), 3.5);
        int v = (int)__dadd_rn(n, 0.5);  // C truncation toward zero
        v = v < 0 ? 0 : (v > 7 ? 7 : v);
        q[deinterleave_inv(j)] = (uint8_t)v;
    }
    __syncwarp();

    // ---- per-step table offsets: (sg1 | sg2 << 3) * 8, in place over the symbol pairs
    uint16_t* qoff = reinterpret_cast<uint16_t*>(q);
    for (int t = lane; t < kFrameBits; t += 32) {
        const uint32_t sg = qoff[t];  // sg1 | sg2 << 8
        qoff[t] = (uint16_t)(((sg & 7u) | ((sg >> 5) & 0x38u)) << 3);
    }
    __syncwarp();

    // ---- forward pass
    const uint32_t k1 = __popc(lane & 0x4F) & 1 ? 7u : 0u;  // parity(j & G1) -> expected coded bit 1
    const uint32_t k2 = __popc(lane & 0x6D) & 1 ? 7u : 0u;
    const uint32_t xmask = (k1 | (k2 << 3)) << 3;            // table offset XOR of this lane
    const uint32_t sel = (lane & 1) ? 0x3232u : 0x1010u;     // which half of the fetched pairs this lane needs
    const int src0 = lane >> 1, src1 = 16 + (lane >> 1);
    const unsigned char* tblb = reinterpret_cast<const unsigned char*>(bm_tbl);
    // aa = a | a << 16, bb = b | b << 16 with a = metric[state lane], b = metric[state lane + 32]
    uint32_t aa = lane == 0 ? 0u : (uint32_t)kUnreachable * 0x00010001u;
    uint32_t bb = (uint32_t)kUnreachable * 0x00010001u;
    uint32_t dreg = 0;
    uint32_t newp = 0;
#pragma unroll 16
    for (int t = 0; t < kFrameBits; ++t) {
        const uint32_t off = qoff[t] ^ xmask;  // lane-uniform load
        const uint2 bm = *reinterpret_cast<const uint2*>(tblb + off);
        const uint32_t X = aa + bm.x, Y = bb + bm.y;  // lo halves: state 2j (input 0), hi halves: state 2j+1
        newp = __vmins2(X, Y);
        // per half Y + 0x8000 - X stays inside 16 bits (metrics < 2^15): bit 15 set <=> X <= Y <=> predecessor p0
        const uint32_t D = (Y + 0x80008000u) - X;
        dreg = (dreg >> 1) | (~D & 0x80008000u);       // decision 1 = came from p1; after 16 steps bit k / 16 + k
        if ((t & 15) == 15) {
            dec[(t >> 4) * 32 + lane] = dreg;
            dreg = 0;
        }
        // redistribute: lane j needs new metrics of states j (lane j>>1) and j+32 (lane 16 + j>>1)
        const uint32_t v0 = __shfl_sync(0xffffffffu, newp, src0);
        const uint32_t v1 = __shfl_sync(0xffffffffu, newp, src1);
        aa = __byte_perm(v0, 0, sel);
        bb = __byte_perm(v1, 0, sel);
    }
    __syncwarp();

    // ---- best final state: lowest metric, lowest index on ties (:835-837)
    const uint32_t m_lo = newp & 0xFFFFu, m_hi = newp >> 16;
    uint32_t key = min((m_lo << 8) | (uint32_t)(2 * lane), (m_hi << 8) | (uint32_t)(2 * lane + 1));
    key = __reduce_min_sync(0xffffffffu, key);

    // ---- traceback + pack + derandomise (:839-843, :878-895)
    // The state register is kept left-aligned in S (state = S >> 26) and doubles as the shift register of the
    // decisions: a step is S = (d : S) >> 1 (one funnel shift), the bit the reference emits at that step
    // (state & 1, :840) is bit 26 before the shift, so after the 8 steps of an output byte the byte is
    // (S >> 18) & 0xFF with the first-emitted bit in bit 0 (:878-884).  7 instructions per step on lane 0.
    if (lane == 0) {
        uint32_t S = (key & 0xFFu) << 26;
        const unsigned char* decb = reinterpret_cast<const unsigned char*>(dec);
#pragma unroll 1
        for (int wd = kDecWords - 1; wd >= 0; --wd) {
            const unsigned char* wb = decb + wd * 128;  // decision words of steps 16 wd .. 16 wd + 15, one per lane
#pragma unroll
            for (int k = 15; k >= 0; --k) {             // step t = 16 wd + k
                const uint32_t w = *reinterpret_cast<const uint32_t*>(wb + ((S >> 25) & 0x7Cu));  // lane state >> 1
                const uint32_t hi = w >> (((S >> 22) & 16u) | (uint32_t)k);  // bit k (even state) or 16 + k (odd)
                S = __funnelshift_r(S, hi, 1);                               // state = state >> 1 | d << 5 (:841-842)
                if (k == 8 || k == 0) {
                    const int i = 2 * (kDecWords - 1 - wd) + (k == 0 ? 1 : 0);
                    outb[i] = (uint8_t)((S >> 18) ^ c_lfsr[i]);
                }
            }
        }
        const int metric = (int)(key >> 8);
        *out_metric = metric;
        atomicAdd(&counters[kCtrFramesDecoded], 1ull);
        if (metric == 0) atomicAdd(&counters[kCtrFramesPerfect], 1ull);
        atomicAdd(&counters[kCtrAcs], (unsigned long long)kFrameBits * 64ull);
    }
    __syncwarp();
    for (int i = lane; i < kFrameBytes; i += 32) out_frame[i] = outb[i];
}

__global__ void __launch_bounds__(32 * kDecWarps)
decode_tasks_kernel(SoftBuffers so, const FrameTask* __restrict__ tasks, const int32_t* __restrict__ n_tasks_dev,
                    int max_tasks, uint8_t* __restrict__ frames, int32_t* __restrict__ metrics, int max_frames,
                    const FrameRec* __restrict__ frec, FrameLogEntry* __restrict__ log, const unsigned long long* log_count,
                    long long log_cap, unsigned long long* __restrict__ counters) {
    extern __shared__ __align__(16) unsigned char dsm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int n = *n_tasks_dev;
    if (n > max_tasks) n = max_tasks;
    const unsigned long long log_base = *log_count;  // frames logged by earlier runs (advanced after this kernel)
    uint2* bm_tbl = reinterpret_cast<uint2*>(dsm);
    build_bm_table(bm_tbl);
    unsigned char* wsm = dsm + kBmTableBytes + (size_t)warp * kDecSmemPerWarp;
    const long long wrap = so.ring ? so.stride : (1ll << 62);
    for (int task = blockIdx.x * kDecWarps + warp; task < n; task += gridDim.x * kDecWarps) {
        const FrameTask ft = tasks[task];
        const double* soft_row = so.soft + (long long)ft.stream * so.stride;
        const long long o = (long long)ft.stream * max_frames + (ft.slot % max_frames);
        decode_one(soft_row, so.ring ? ft.payload_start % so.stride : ft.payload_start, wrap, wsm, bm_tbl,
                   frames + o * kFrameBytes, metrics + o, counters);
        __syncwarp();
        // ---- the same frame into the contiguous log the host polls (new entries only cross PCIe)
        FrameLogEntry* e = log + (long long)((log_base + (unsigned long long)task) % (unsigned long long)log_cap);
        const uint8_t* outb = wsm + kDecBytesPerWarp;  // decode_one leaves the 134 bytes here (zeros when dropped)
        int metric = 0;
        if (lane == 0) {
            metric = metrics[o];  // written by this lane in decode_one
            e->stream = ft.stream; e->frame_idx = ft.slot; e->metric = metric; e->reserved = 0;
            e->payload_start = ft.payload_start; e->ready_idx = frec[o].ready_idx; e->quality = frec[o].quality;
        }
        metric = __shfl_sync(0xffffffffu, metric, 0);
        if (metric >= 0)
            for (int i = lane; i < kFrameBytes; i += 32) e->frame[i] = outb[i];
        __syncwarp();
    }
}

// after the decoder: the frames of this run become visible to the host's poll
__global__ void log_advance_kernel(unsigned long long* log_count, const int32_t* n_tasks_dev, int max_tasks) {
    int n = *n_tasks_dev;
    if (n > max_tasks) n = max_tasks;
    *log_count += (unsigned long long)n;
}

__global__ void __launch_bounds__(32 * kDecWarps)
decode_payloads_kernel(const double* __restrict__ payloads, int n, uint8_t* __restrict__ frames,
                       int32_t* __restrict__ metrics, unsigned long long* __restrict__ counters) {
    extern __shared__ __align__(16) unsigned char dsm[];
    const int warp = threadIdx.x >> 5;
    uint2* bm_tbl = reinterpret_cast<uint2*>(dsm);
    build_bm_table(bm_tbl);
    unsigned char* wsm = dsm + kBmTableBytes + (size_t)warp * kDecSmemPerWarp;
    for (int task = blockIdx.x * kDecWarps + warp; task < n; task += gridDim.x * kDecWarps) {
        decode_one(payloads + (long long)task * kEncodedBits, 0, 1ll << 62, wsm, bm_tbl,
                   frames + (long long)task * kFrameBytes, metrics + task, counters);
        __syncwarp();
    }
}

static int decode_grid(int n_tasks) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int per_sm = 10;  // 10 CTAs x 2 warps x 10.9 KB = 217 KB of shared memory per SM
    int want = (n_tasks + kDecWarps - 1) / kDecWarps;
    int cap = sms * per_sm;
    return want < cap ? (want > 0 ? want : 1) : cap;
}

void launch_decode(const SoftBuffers& so, const FrameTask* tasks, const int32_t* n_tasks_dev, int n_tasks_host,
                   uint8_t* frames, int32_t* metrics, int max_frames, const FrameRec* frec, FrameLogEntry* log,
                   unsigned long long* log_count, long long log_cap, unsigned long long* counters, cudaStream_t st) {
    // n_tasks_host is an upper bound used only to size the grid; the kernel reads the exact count on device
    if (n_tasks_host <= 0) return;
    const size_t smem = (size_t)kBmTableBytes + (size_t)kDecWarps * kDecSmemPerWarp;
    cudaFuncSetAttribute(decode_tasks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  // per device
    prefer_max_shared(decode_tasks_kernel);
    prefer_max_shared(log_advance_kernel);
    decode_tasks_kernel<<<decode_grid(n_tasks_host), 32 * kDecWarps, smem, st>>>(so, tasks, n_tasks_dev, n_tasks_host,
                                                                              frames, metrics, max_frames, frec, log,
                                                                              log_count, log_cap, counters);
    log_advance_kernel<<<1, 1, 0, st>>>(log_count, n_tasks_dev, n_tasks_host);
}

void launch_decode_payloads(const double* payloads, int n, uint8_t* frames, int32_t* metrics,
                            unsigned long long* counters, cudaStream_t st) {
    if (n <= 0) return;
    const size_t smem = (size_t)kBmTableBytes + (size_t)kDecWarps * kDecSmemPerWarp;
    cudaFuncSetAttribute(decode_payloads_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  // per device
    prefer_max_shared(decode_payloads_kernel);
    decode_payloads_kernel<<<decode_grid(n), 32 * kDecWarps, smem, st>>>(payloads, n, frames, metrics, counters);
}

}  // namespace opvd
