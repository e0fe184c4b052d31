// opvd_kernels.cuh — launch interface between the host runtime (opvd_api.cu) and the sm_100a
// kernels (kernels_*.cu).  Everything here is plain device pointers and sizes.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#include "demod_core.cuh"
#include "track_core.cuh"

namespace opvd {

// One decoded-frame work item produced by the tracker and consumed by the decoder.
struct FrameTask {
    int32_t stream;
    int32_t slot;           // frame index within the stream (frame_ready order)
    int64_t payload_start;  // absolute symbol index of the first payload soft symbol
};

// One decoded frame in the contiguous log the host polls: the decoder appends the frames of every run in task order
// (in order within a stream), so a poll copies only what is new.  176 bytes.
struct FrameLogEntry {
    int32_t stream, frame_idx, metric, reserved;
    int64_t payload_start, ready_idx;
    double quality;
    uint8_t frame[kFrameBytes];
    uint8_t pad[2];
};

// device-side counters (uint64 each); reduced across ranks by the caller
enum Counter : int {
    kCtrSamples = 0,     // samples consumed by the demodulator (call origins advanced + last call)
    kCtrSymbols,         // soft symbols produced
    kCtrFramesReady,     // tracker frame_ready events
    kCtrFramesDecoded,   // metric >= 0
    kCtrFramesPerfect,   // metric == 0
    kCtrFramesDropped,   // scale < 1e-10
    kCtrSyncAcq,         // HUNTING -> VERIFYING
    kCtrSyncOk,
    kCtrSyncMiss,
    kCtrLostLock,
    kCtrBitErrors,       // vs known BERT payloads (filled by bert_check)
    kCtrFramesCompared,
    kCtrAcs,             // Viterbi add-compare-select state updates
    kNumCounters = 16
};

// Per-stream sample rows.  Two layouts:
//   linear (ring == 0): row[x] holds absolute sample x of the stream (attached captures, batch mode);
//   ring   (ring == 1): row[x % stride] holds absolute sample x (library-owned buffer in stream mode): the host
//                       appends behind `avail` and never moves anything; what a launch may still read is
//                       [origin - 192, avail) of each stream, and the host keeps that span intact (opvd_api.cu).
struct StreamBuffers {
    const uint32_t* iq;      // packed int16 I/Q, row-major [stream][stride]
    int64_t stride;          // samples per row (multiple of 4; multiple of 64 for a ring)
    const int64_t* avail;    // [S] absolute sample count pushed so far per stream
    int32_t ring;
};

// Soft symbols, same two layouts over absolute symbol indices.
struct SoftBuffers {
    double* soft;            // [stream][stride]
    int64_t stride;
    int32_t ring;
    int64_t* n_sym;          // [S] symbols available after this launch: written by the demodulator, read by the
                             // tracker (its own array per run in flight, so tile t+1 may be demodulated while
                             // tile t is still being tracked and decoded)
};

#if defined(__CUDACC__)
// One stream's row as seen by one launch: rel = x - base_abs is a small int, the physical offset is
// base_off + rel, minus the ring length when it wraps (never for a linear row).
struct RowView {
    const uint32_t* row;
    long long base_abs;      // absolute sample index of rel 0 (multiple of 64, <= origin - 128 or 0)
    int base_off;            // physical offset of rel 0
    int wrap;                // ring: stride; linear: INT_MAX
    int rel_end;             // linear: rel values >= rel_end lie beyond the row; ring: INT_MAX
    __device__ __forceinline__ int phys(int rel) const {
        int o = base_off + rel;
        if (o >= wrap) o -= wrap;
        if (o >= wrap) o -= wrap;
        return o;
    }
};
__device__ __forceinline__ RowView make_row_view(const StreamBuffers& sb, int stream, long long origin) {
    RowView v;
    v.row = sb.iq + (long long)stream * sb.stride;
    long long b = origin - 128;
    v.base_abs = b < 0 ? 0 : (b & ~63ll);
    if (sb.ring) {
        v.base_off = (int)(v.base_abs % sb.stride);
        v.wrap = (int)sb.stride;
        v.rel_end = 0x7fffffff;
    } else {
        v.base_off = (int)v.base_abs;
        v.wrap = 0x7fffffff;
        const long long e = sb.stride - v.base_abs;
        v.rel_end = e > 0x7fffffff ? 0x7fffffff : (int)e;
    }
    return v;
}
// soft-symbol row position of absolute symbol n, and its successor
__device__ __forceinline__ long long soft_pos(const SoftBuffers& so, long long n) { return so.ring ? n % so.stride : n; }
#endif

void launch_estimate(const StreamBuffers& sb, DemodState* dstate, double* est_out, int n_streams, int mode,
                     int final_flag, cudaStream_t st);

// lanes_per_stream in {32, 96}; 0 = chosen from the stream count (demod_select.cu); returns cudaError
//   32   warp per stream (kernels_demod_warp.cu)      96   channel bank, three role warps per 32 streams
// Every kernel of the chain asks for the same (largest) shared-memory carveout: an SM can only change its L1/shared
// split while it is empty, so a kernel with a different preference cannot join CTAs that are already resident - the
// tracker and the Viterbi decoder of tile t then wait for the demodulator of tile t+1 to drain instead of running
// beside it (measured with OPVD_TRACE: every second decode launch took a whole tile).
template <class Kernel>
inline void prefer_max_shared(Kernel kernel) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
}

cudaError_t launch_demod(const StreamBuffers& sb, const SoftBuffers& so, DemodState* dstate, int n_streams,
                         int mode, int final_flag, double afc_alpha, int lanes_per_stream,
                         unsigned long long* counters, cudaStream_t st);

// the lanes_per_stream value launch_demod resolves 0 (automatic) to
int demod_auto_lanes(int n_streams);

// warp-per-stream variant (kernels_demod_warp.cu); selected by launch_demod for lanes_per_stream == 32
cudaError_t launch_demod_warp(const StreamBuffers& sb, const SoftBuffers& so, DemodState* dstate, int n_streams,
                              int mode, int final_flag, double afc_alpha, unsigned long long* counters,
                              cudaStream_t st);

// channel-bank variant (kernels_demod_bank.cu): 32 streams per 96-thread CTA, three free-running role warps;
// selected by launch_demod for lanes_per_stream == 96
cudaError_t launch_demod_bank(const StreamBuffers& sb, const SoftBuffers& so, DemodState* dstate, int n_streams,
                              int mode, int final_flag, double afc_alpha, unsigned long long* counters,
                              cudaStream_t st);

// coherent mode (kernels_demod_coherent.cu): `opv-demod -c`, batch only
cudaError_t launch_demod_coherent(const StreamBuffers& sb, const SoftBuffers& so, DemodState* dstate, int n_streams,
                                  int final_flag, double afc_alpha, double pll_bw_hz, unsigned long long* counters,
                                  cudaStream_t st);

void launch_track(const SoftBuffers& so, TrackState* tstate, int n_streams,
                  FrameRec* frec, int max_frames, TrackEvent* events, int32_t* n_events, int max_events,
                  FrameTask* tasks, int32_t* n_tasks, int max_tasks, unsigned long long* counters,
                  cudaStream_t st);

void launch_decode(const SoftBuffers& so, const FrameTask* tasks, const int32_t* n_tasks_dev, int n_tasks_host,
                   uint8_t* frames, int32_t* metrics, int max_frames, const FrameRec* frec, FrameLogEntry* log,
                   unsigned long long* log_count, long long log_cap, unsigned long long* counters, cudaStream_t st);

// stage-level entry (FrameDecoder::decode seam): payloads [n][2144] doubles -> frames [n][134], metrics [n]
void launch_decode_payloads(const double* payloads, int n, uint8_t* frames, int32_t* metrics,
                            unsigned long long* counters, cudaStream_t st);

// synthetic OPV bank generator (measurement aid; opv-mod-like TX chain + impairments)
struct SynthParams {
    int n_streams;
    int n_frames;            // frames per stream
    int64_t stride;          // samples per row
    int64_t n_samples;       // samples generated per row (lead + n_frames*86720 + tail)
    uint64_t seed;
    float scale;             // headroom scaling (0.25)
    float ebn0_lo_db, ebn0_hi_db;   // per-stream Eb/N0 swept linearly across streams; <= -100 disables noise
    float cfo_max_hz;        // per-stream CFO uniform in [-cfo_max, +cfo_max]
    int frac_delay;          // 1: per-stream fractional delay in [0,1)
    int max_lead;            // per-stream leading gap uniform in [0, max_lead) samples (noise only)
    int first_stream;        // global index of stream 0 of this rank (seeds are global)
};
void launch_synth(const SynthParams& p, uint32_t* iq, uint8_t* scratch_syms, int8_t* scratch_sign,
                  cudaStream_t st);
size_t synth_scratch_bytes(const SynthParams& p);

void launch_bert_check_impl(const uint8_t* frames, const int32_t* metrics, const FrameRec* frec,
                            const int32_t* n_frames_per_stream, int n_streams, int max_frames, const SynthParams& p,
                            unsigned long long* counters, cudaStream_t st);

void upload_constants();

}  // namespace opvd
