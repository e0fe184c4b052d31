// kernels_track.cu — A5: sync correlator + flywheel, one warp per stream.
//
// The reference walks every soft symbol through SyncTracker::process
// (/root/reference/src/opv-demod.cpp:615-736).  Here the state machine is advanced by EVENTS
// (track_core.cuh): while LOCKED / VERIFYING nothing happens between frame boundaries, so the
// warp jumps 2168 symbols at a time; while HUNTING the 24-tap soft correlation is a feed-forward
// sliding window, so the 32 lanes test 32 consecutive symbol positions at once (each lane keeps
// the reference's sequential summation order) and a ballot picks the first hit.
#include <cuda_runtime.h>
#include <cstdint>

#include "opvd_kernels.cuh"

namespace opvd {

constexpr int kTrackWarps = 2;  // small CTAs: they fit beside resident demodulator CTAs (overlap with the next tile)

__global__ void __launch_bounds__(32 * kTrackWarps)
track_kernel(SoftBuffers so, TrackState* __restrict__ tstate, int n_streams,
             FrameRec* __restrict__ frec, int max_frames, TrackEvent* __restrict__ events,
             int32_t* __restrict__ n_events, int max_events, FrameTask* __restrict__ tasks,
             int32_t* __restrict__ n_tasks, int max_tasks, unsigned long long* __restrict__ counters) {
    const int lane = threadIdx.x & 31;
    const int stream = blockIdx.x * kTrackWarps + (threadIdx.x >> 5);
    if (stream >= n_streams) return;

    TrackState t = tstate[stream];
    const long long N = so.n_sym[stream];  // soft symbols available (this run's snapshot, written by the demodulator)
    const double* soft_row = so.soft + (long long)stream * so.stride;
    const long long wrap = so.ring ? so.stride : 0x7fffffffffffffffll;
    // the 24 soft symbols ending at absolute symbol n, in order (the row may be a ring)
    auto correlate_at = [&](long long n, double& raw) {
        long long p = n - (kSyncBits - 1);
        if (so.ring) p %= so.stride;
        if (p + kSyncBits <= wrap) return sync_correlate(soft_row + p, raw);
        double w[kSyncBits];
#pragma unroll
        for (int i = 0; i < kSyncBits; ++i) {
            long long q = p + i;
            if (q >= wrap) q -= wrap;
            w[i] = soft_row[q];
        }
        return sync_correlate(w, raw);
    };
    int ne = n_events[stream];
    unsigned long long c_ready = 0, c_acq = 0, c_ok = 0, c_miss = 0, c_lost = 0;

    auto push_event = [&](int type, int count, long long idx, double corr, double raw) {
        if (lane == 0) {
            TrackEvent e;
            e.type = type; e.count = count; e.sym_idx = idx; e.corr = corr; e.raw = raw;
            events[(long long)stream * max_events + (ne % max_events)] = e;
        }
        ++ne;  // the log is a ring: the host reads the new entries after every run
    };
    auto push_frame = [&](long long payload_start, long long ready, double quality) {
        const int slot = t.total_frames;
        if (lane == 0) {
            {
                FrameRec fr;  // per-stream ring of max_frames slots, drained by the host after every run
                fr.payload_start = payload_start; fr.ready_idx = ready; fr.quality = quality;
                frec[(long long)stream * max_frames + (slot % max_frames)] = fr;
                const int k = atomicAdd(n_tasks, 1);
                if (k < max_tasks) {
                    FrameTask ft;
                    ft.stream = stream; ft.slot = slot; ft.payload_start = payload_start;
                    tasks[k] = ft;
                }
            }
        }
        t.total_frames++;
        ++c_ready;
    };

    for (;;) {
        if (t.state == kHunting) {
            long long n0 = t.cursor < (kSyncBits - 1) ? (kSyncBits - 1) : t.cursor;
            bool found = false;
            while (n0 < N) {
                const long long n = n0 + lane;
                double raw = 0.0, norm = 0.0;
                bool hit = false;
                if (n < N) {
                    norm = correlate_at(n, raw);
                    hit = hunt_hit(norm, raw);
                }
                const unsigned m = __ballot_sync(0xffffffffu, hit);
                if (m) {
                    const int src = __ffs(m) - 1;
                    const double hn = __shfl_sync(0xffffffffu, norm, src);
                    const double hr = __shfl_sync(0xffffffffu, raw, src);
                    const long long nh = n0 + src;
                    t.state = kVerifying; t.quality = hn; t.anchor = nh; t.collecting = 1; t.payload_start = nh + 1;
                    push_event(kEvHuntToVerify, 0, nh, hn, hr);
                    ++c_acq;
                    found = true;
                    break;
                }
                n0 += 32;
            }
            if (!found) { t.cursor = N; break; }
        } else if (t.state == kVerifying) {
            const long long ready = t.anchor + kEncodedBits;  // :658
            if (ready >= N) break;
            push_frame(t.payload_start, ready, t.quality);
            t.collecting = 0; t.state = kLocked; t.misses = 0;
            push_event(kEvVerifyToLocked, t.total_frames, ready, 0.0, 0.0);
        } else {  // LOCKED
            if (t.collecting) {
                const long long ready = t.payload_start + kEncodedBits - 1;  // :720
                if (ready >= N) break;
                push_frame(t.payload_start, ready, t.quality);
                t.collecting = 0;
            }
            const long long nb = t.anchor + kFrameSymbols;  // :684
            if (nb >= N) break;
            double raw;
            const double corr = correlate_at(nb, raw);  // lane-uniform
            if (corr >= 0.70) {  // :688
                t.misses = 0; t.quality = corr; t.collecting = 1; t.payload_start = nb + 1; t.anchor = nb;
                push_event(kEvSyncOk, 0, nb, corr, raw);
                ++c_ok;
            } else {
                t.misses++;
                push_event(kEvSyncMiss, t.misses, nb, corr, raw);
                ++c_miss;
                if (t.misses >= kSyncMissLimit) {  // :702-707
                    t.state = kHunting; t.collecting = 0; t.cursor = nb + 1;
                    push_event(kEvLostLock, 0, nb, 0.0, 0.0);
                    ++c_lost;
                } else {  // flywheel :709-712
                    t.quality = corr; t.collecting = 1; t.payload_start = nb + 1; t.anchor = nb;
                }
            }
        }
    }
    if (lane == 0) {
        tstate[stream] = t;
        n_events[stream] = ne;
        if (c_ready) atomicAdd(&counters[kCtrFramesReady], c_ready);
        if (c_acq) atomicAdd(&counters[kCtrSyncAcq], c_acq);
        if (c_ok) atomicAdd(&counters[kCtrSyncOk], c_ok);
        if (c_miss) atomicAdd(&counters[kCtrSyncMiss], c_miss);
        if (c_lost) atomicAdd(&counters[kCtrLostLock], c_lost);
    }
}

void launch_track(const SoftBuffers& so, TrackState* tstate, int n_streams,
                  FrameRec* frec, int max_frames, TrackEvent* events, int32_t* n_events, int max_events,
                  FrameTask* tasks, int32_t* n_tasks, int max_tasks, unsigned long long* counters,
                  cudaStream_t st) {
    if (n_streams <= 0) return;
    const int grid = (n_streams + kTrackWarps - 1) / kTrackWarps;
    prefer_max_shared(track_kernel);
    track_kernel<<<grid, 32 * kTrackWarps, 0, st>>>(so, tstate, n_streams, frec, max_frames, events,
                                                    n_events, max_events, tasks, n_tasks, max_tasks, counters);
}

}  // namespace opvd
