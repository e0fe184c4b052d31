// kernels_demod_coherent.cu — the reference's batch-only coherent demodulator (`opv-demod -c`,
// CoherentMSKDemodulator, /root/reference/src/opv-demod.cpp:365-572, selected at :1144-1161) for
// sm_100a; arithmetic in demod_coherent_core.cuh.  SURVEY section 8(f) rank 3: completes the CLI.
//
// Not a throughput path: the reference's Costas loop does not lock on Opulent Voice captures (see
// DESIGN.md), so this kernel exists for contract completeness.  One lane per stream; the symbol grid
// is fixed (40 samples, no timing recovery), so every lane streams its own row with 128-bit loads
// (symbol k starts at byte 160 k: always 16-byte aligned) and needs no shared memory.
#include <cuda_runtime.h>
#include <cstdint>

#include "demod_coherent_core.cuh"
#include "opvd_kernels.cuh"

namespace opvd {

__global__ void __launch_bounds__(64)
demod_coherent_kernel(StreamBuffers sb, SoftBuffers so, DemodState* __restrict__ dstate, int n_streams, int final_flag,
                      double afc_alpha, double pll_bw_hz, unsigned long long* __restrict__ counters) {
    const int stream = blockIdx.x * blockDim.x + threadIdx.x;
    if (stream >= n_streams || !final_flag) return;  // batch mode: load everything, then process (:1132-1135)
    DemodState st = dstate[stream];
    if (st.flags & kFlagDone) return;
    const long long avail = sb.avail[stream];
    const uint4* row = reinterpret_cast<const uint4*>(sb.iq + (long long)stream * sb.stride);  // batch mode: always a linear row
    double* soft_row = so.soft + (long long)stream * so.stride;
    CoherentState cs;
    coherent_init(cs, st.freq_offset, afc_alpha, pll_bw_hz);  // freq_offset = estimate (:1148-1149)
    const long long n_sym = avail / kSps;
    for (long long sym = 0; sym < n_sym; ++sym) {
        double I[kSps], Q[kSps];
#pragma unroll
        for (int v = 0; v < kSps / 4; ++v) {
            const uint4 w = __ldg(row + sym * (kSps / 4) + v);
            unpack_iq(w.x, I[4 * v + 0], Q[4 * v + 0]);
            unpack_iq(w.y, I[4 * v + 1], Q[4 * v + 1]);
            unpack_iq(w.z, I[4 * v + 2], Q[4 * v + 2]);
            unpack_iq(w.w, I[4 * v + 3], Q[4 * v + 3]);
        }
        soft_row[sym] = coherent_symbol(cs, I, Q, sym == 0);
    }
    st.freq_offset = cs.freq_offset;
    st.n_sym = n_sym;
    st.origin = avail;
    st.flags |= kFlagDone;
    dstate[stream] = st;
    so.n_sym[stream] = n_sym;
    if (n_sym) atomicAdd(&counters[kCtrSymbols], (unsigned long long)n_sym);
    if (avail) atomicAdd(&counters[kCtrSamples], (unsigned long long)avail);
}

cudaError_t launch_demod_coherent(const StreamBuffers& sb, const SoftBuffers& so, DemodState* dstate, int n_streams,
                                  int final_flag, double afc_alpha, double pll_bw_hz, unsigned long long* counters,
                                  cudaStream_t st) {
    if (n_streams <= 0) return cudaSuccess;
    const int threads = 64;
    demod_coherent_kernel<<<(n_streams + threads - 1) / threads, threads, 0, st>>>(sb, so, dstate, n_streams, final_flag,
                                                                                  afc_alpha, pll_bw_hz, counters);
    return cudaGetLastError();
}

}  // namespace opvd
