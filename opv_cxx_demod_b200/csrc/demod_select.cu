// demod_select.cu — which demodulator kernel runs for a bank (A1/A3/A4).
//
// A stream is a strictly serial recurrence (symbol n+1's window position and LO step depend on symbol n), so the
// parallel axes are streams and the inside of one symbol, and the fastest organisation depends on how many streams
// there are per SM:
//   lanes_per_stream = 32   one warp per stream (kernels_demod_warp.cu): low per-symbol latency, for small banks
//   lanes_per_stream = 96   channel-bank kernel, three role warps per 32 streams (kernels_demod_bank.cu): large banks
#include <cuda_runtime.h>

#include "opvd_kernels.cuh"

namespace opvd {

int demod_auto_lanes(int n_streams) {
    // Measured crossover (profiles/gpu_r02_wb.log, gpu_r02_elb.log, 12-frame banks): the warp-per-stream kernel holds 8
    // CTAs per SM in its fast build (1,184 streams, 16.2 ms at 1,024) and 16 in its 124-register build (23.3 ms at 1,536,
    // 29.7 ms at 2,048); the bank kernel needs 28.5 ms for anything up to one CTA per SM (4,736 streams).
    int sms = 148, dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if ((long long)n_streams <= 12ll * sms) return 32;
    return 96;
}

cudaError_t launch_demod(const StreamBuffers& sb, const SoftBuffers& so, DemodState* dstate, int n_streams,
                         int mode, int final_flag, double afc_alpha, int lanes_per_stream,
                         unsigned long long* counters, cudaStream_t st) {
    if (n_streams <= 0) return cudaSuccess;
    const int L = lanes_per_stream > 0 ? lanes_per_stream : demod_auto_lanes(n_streams);
    if (L >= 96) return launch_demod_bank(sb, so, dstate, n_streams, mode, final_flag, afc_alpha, counters, st);
    return launch_demod_warp(sb, so, dstate, n_streams, mode, final_flag, afc_alpha, counters, st);
}

}  // namespace opvd
