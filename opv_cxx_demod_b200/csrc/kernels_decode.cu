// kernels_decode.cu — A6/A7/A8: frame decoder, one warp per frame.
//   scale + 3-bit quantise           FrameDecoder::decode  /root/reference/src/opv-demod.cpp:856-866
//   67x32 bit-reversed deinterleave  deinterleave_addr     :792-795   (fused into the Viterbi input load)
//   K=7 r=1/2 Viterbi                ViterbiDecoder::decode :802-846
//   pack + LFSR derandomise          :878-895             (fused into the traceback epilogue)
//
// Viterbi layout: 64 states, 2 per lane.  Lane j holds the path metrics of states j and j+32
// (the two predecessors of states 2j and 2j+1), packed as 2 x int16 in one register, so the
// add-compare-select butterfly is lane-local: two 32-bit adds form all four candidates, one DPX
// __vibmin_s16x2 does both compare-selects and returns the two decision predicates with the
// reference's tie rule (a <= b keeps predecessor p0, :829).  The new metrics are redistributed with
// two warp shuffles + one byte-permute.  Metrics are bounded by 2144*7 = 15,008, so int16 holds
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

constexpr int kDecWarps = 4;
constexpr int kDecWords = kFrameBits / 16;                 // 67 decision words per lane
constexpr int kDecBytesPerWarp = kDecWords * 32 * 4;       // 8576 B, aliased as 1072 doubles for the scale sum
constexpr int kQBytesPerWarp = kEncodedBits;               // deinterleaved 3-bit symbols, one byte each
constexpr int kOutBytesPerWarp = 144;
constexpr int kDecSmemPerWarp = kDecBytesPerWarp + kQBytesPerWarp + kOutBytesPerWarp;  // 10,864 B
constexpr int kUnreachable = 16384;

// inverse of deinterleave_addr: deint[i] = q[deinterleave_addr(i)]  =>  q index j feeds deint index inv(j)
__device__ __forceinline__ int deinterleave_inv(int j) {
    const int pos = (j & ~7) + 7 - (j & 7);
    return (pos % 67) * 32 + pos / 67;
}

__device__ __forceinline__ void decode_one(const double* __restrict__ soft, unsigned char* wsm, uint8_t* out_frame,
                                           int32_t* out_metric, unsigned long long* counters) {
    const int lane = threadIdx.x & 31;
    double* dsum = reinterpret_cast<double*>(wsm);
    uint32_t* dec = reinterpret_cast<uint32_t*>(wsm);
    uint8_t* q = wsm + kDecBytesPerWarp;
    uint8_t* outb = q + kQBytesPerWarp;

    // ---- scale = mean |soft| with the reference's sequential summation order (:856-858)
    double scale = 0.0;
    for (int half = 0; half < 2; ++half) {
        const double* src = soft + half * (kEncodedBits / 2);
        for (int i = lane; i < kEncodedBits / 2; i += 32) dsum[i] = fabs(src[i]);
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
        const double n = __dadd_rn(__dmul_rn(__ddiv_rn(-soft[j], scale), 3.5), 3.5);
        int v = (int)__dadd_rn(n, 0.5);  // C truncation toward zero
        v = v < 0 ? 0 : (v > 7 ? 7 : v);
        q[deinterleave_inv(j)] = (uint8_t)v;
    }
    __syncwarp();

    // ---- forward pass
    const uint32_t k1 = __popc(lane & 0x4F) & 1 ? 7u : 0u;  // parity(j & G1) -> expected coded bit 1
    const uint32_t k2 = __popc(lane & 0x6D) & 1 ? 7u : 0u;
    const uint32_t psel = (lane & 1) ? 0x7632u : 0x5410u;
    // ab = metric[state lane] | metric[state lane+32] << 16
    uint32_t ab = (lane == 0 ? 0u : (uint32_t)kUnreachable) | ((uint32_t)kUnreachable << 16);
    uint32_t dreg = 0;
    uint32_t newp = 0;
    const uint16_t* q2 = reinterpret_cast<const uint16_t*>(q);
    for (int t = 0; t < kFrameBits; ++t) {
        const uint32_t sg = q2[t];  // sg1 | sg2 << 8, lane-uniform
        const uint32_t u1 = (sg & 0xFFu) ^ k1;   // cost of coded bit e1 on predecessor lane (in=0)
        const uint32_t u2 = (sg >> 8) ^ k2;
        const uint32_t A0 = u1 + u2;             // p0=state j,    in=0
        const uint32_t B0 = u1 + (u2 ^ 7u);      // p1=state j+32, in=0 (G2 taps bit 5, G1 does not)
        // in=1 flips both coded bits: A1 = 14 - A0, B1 = 14 - B0
        const uint32_t bmA = A0 * 0xFFFF0001u + 0x000E0000u;  // A0 | (14-A0) << 16
        const uint32_t bmB = B0 * 0xFFFF0001u + 0x000E0000u;
        const uint32_t aa = __byte_perm(ab, 0, 0x1010);  // a | a << 16
        const uint32_t bb = __byte_perm(ab, 0, 0x3232);  // b | b << 16
        bool p_hi, p_lo;
        newp = __vibmin_s16x2(aa + bmA, bb + bmB, &p_hi, &p_lo);  // lo: state 2j, hi: state 2j+1
        const uint32_t d = (p_lo ? 0u : 1u) | (p_hi ? 0u : 2u);    // decision 1 = came from p1
        dreg |= d << (2 * (t & 15));
        if ((t & 15) == 15) {
            dec[(t >> 4) * 32 + lane] = dreg;
            dreg = 0;
        }
        // redistribute: lane j needs new metrics of states j (lane j>>1) and j+32 (lane 16 + j>>1)
        const uint32_t v0 = __shfl_sync(0xffffffffu, newp, lane >> 1);
        const uint32_t v1 = __shfl_sync(0xffffffffu, newp, 16 + (lane >> 1));
        ab = __byte_perm(v0, v1, psel);
    }
    __syncwarp();

    // ---- best final state: lowest metric, lowest index on ties (:835-837)
    const uint32_t m_lo = newp & 0xFFFFu, m_hi = newp >> 16;
    uint32_t key = min((m_lo << 8) | (uint32_t)(2 * lane), (m_hi << 8) | (uint32_t)(2 * lane + 1));
    key = __reduce_min_sync(0xffffffffu, key);

    // ---- traceback + pack + derandomise (:839-843, :878-895)
    if (lane == 0) {
        int s = (int)(key & 0xFFu);
        uint32_t acc = 0;
        for (int t = kFrameBits - 1; t >= 0; --t) {
            const int jb = (kFrameBits - 1 - t) & 7;
            acc |= (uint32_t)(s & 1) << jb;
            const uint32_t w = dec[(t >> 4) * 32 + (s >> 1)];
            const uint32_t d = (w >> (2 * (t & 15) + (s & 1))) & 1u;
            s = (s >> 1) | (int)(d << 5);
            if (jb == 7) {
                const int i = (kFrameBits - 1 - t) >> 3;
                outb[i] = (uint8_t)(acc ^ c_lfsr[i]);
                acc = 0;
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
                    unsigned long long* __restrict__ counters) {
    extern __shared__ __align__(16) unsigned char dsm[];
    const int warp = threadIdx.x >> 5;
    int n = *n_tasks_dev;
    if (n > max_tasks) n = max_tasks;
    unsigned char* wsm = dsm + (size_t)warp * kDecSmemPerWarp;
    for (int task = blockIdx.x * kDecWarps + warp; task < n; task += gridDim.x * kDecWarps) {
        const FrameTask ft = tasks[task];
        const double* soft = so.soft + (long long)ft.stream * so.stride - so.base + ft.payload_start;
        const long long o = (long long)ft.stream * max_frames + (ft.slot % max_frames);
        decode_one(soft, wsm, frames + o * kFrameBytes, metrics + o, counters);
        __syncwarp();
    }
}

__global__ void __launch_bounds__(32 * kDecWarps)
decode_payloads_kernel(const double* __restrict__ payloads, int n, uint8_t* __restrict__ frames,
                       int32_t* __restrict__ metrics, unsigned long long* __restrict__ counters) {
    extern __shared__ __align__(16) unsigned char dsm[];
    const int warp = threadIdx.x >> 5;
    unsigned char* wsm = dsm + (size_t)warp * kDecSmemPerWarp;
    for (int task = blockIdx.x * kDecWarps + warp; task < n; task += gridDim.x * kDecWarps) {
        decode_one(payloads + (long long)task * kEncodedBits, wsm, frames + (long long)task * kFrameBytes,
                   metrics + task, counters);
        __syncwarp();
    }
}

static int decode_grid(int n_tasks) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int per_sm = 5;  // 5 CTAs x 4 warps x 10.9 KB = 217 KB of shared memory per SM
    int want = (n_tasks + kDecWarps - 1) / kDecWarps;
    int cap = sms * per_sm;
    return want < cap ? (want > 0 ? want : 1) : cap;
}

void launch_decode(const SoftBuffers& so, const FrameTask* tasks, const int32_t* n_tasks_dev, int n_tasks_host,
                   uint8_t* frames, int32_t* metrics, int max_frames, unsigned long long* counters,
                   cudaStream_t st) {
    // n_tasks_host is an upper bound used only to size the grid; the kernel reads the exact count on device
    if (n_tasks_host <= 0) return;
    const size_t smem = (size_t)kDecWarps * kDecSmemPerWarp;
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(decode_tasks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(decode_payloads_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr = true;
    }
    decode_tasks_kernel<<<decode_grid(n_tasks_host), 32 * kDecWarps, smem, st>>>(so, tasks, n_tasks_dev, n_tasks_host,
                                                                              frames, metrics, max_frames, counters);
}

void launch_decode_payloads(const double* payloads, int n, uint8_t* frames, int32_t* metrics,
                            unsigned long long* counters, cudaStream_t st) {
    if (n <= 0) return;
    const size_t smem = (size_t)kDecWarps * kDecSmemPerWarp;
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(decode_tasks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(decode_payloads_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr = true;
    }
    decode_payloads_kernel<<<decode_grid(n), 32 * kDecWarps, smem, st>>>(payloads, n, frames, metrics, counters);
}

}  // namespace opvd
