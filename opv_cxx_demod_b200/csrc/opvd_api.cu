// opvd_api.cu — host runtime behind include/opvd.h: owns the device buffers and the per-stream state of the
// receive chain, sequences the sm_100a kernels, and moves frames, events and counters back to caller-owned host
// memory.  No compute happens on the host and there is no CPU fallback: without a CUDA device every entry point
// returns OPVD_ERR_CUDA.
//
// Sustained operation (time tiles, banks larger than HBM) needs no housekeeping pass:
//   * the library-owned sample buffer of a stream-mode handle is a RING per stream (absolute sample x lives at
//     row[x % stride]); opvd_push_iq* appends behind `avail`, nothing is ever moved.  What a run may still read is
//     bounded on the host without reading device state: in stream mode a run leaves less than one chunk
//     (86,720 samples, src/opv-demod.cpp:1012) unconsumed, so everything older than avail(run) - 86,720 - 192 is
//     free once that run's demodulator has finished — the copy stream waits for exactly that event.
//   * soft symbols live in a ring per stream as well; the tracker and the decoder address it modulo its length.
//   * three CUDA streams: copies (H2D of the next tile), front (estimate + demodulate), back (sync tracker +
//     Viterbi + frame log).  back(t) overlaps front(t+1); front(t) waits for back(t-2) (soft ring reuse).
//     opvd_run never synchronises with the device.
//   * decoded frames are appended to a contiguous device log in task order; a poll copies only the new entries.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <deque>
#include <string>
#include <utility>
#include <vector>

#include "../../include/opvd.h"
#include "opvd_kernels.cuh"

using namespace opvd;

static_assert(sizeof(opvd_event) == sizeof(TrackEvent), "event layout");
static_assert(OPVD_NUM_COUNTERS == kNumCounters, "counter layout");
static_assert(sizeof(FrameLogEntry) == 176, "log entry layout");

namespace {
constexpr int kRuns = 8;                          // runs whose bookkeeping (avail snapshot, events) is kept
constexpr int64_t kCarry = kChunkSamples + 192;   // samples a finished stream-mode run may still need (see above)
constexpr double kMinSamplesPerSymbol = 39.895;   // 40 + timing_adj, |timing_adj| <= 0.005 + 0.1 (:283-286)
constexpr int64_t kTrackBack = kFrameSymbols + 2 * kSyncBits;  // symbols the tracker / decoder may reach back
}

struct opvd_handle {
    opvd_config cfg{};
    int dev = 0;
    int S = 0;
    cudaStream_t st = nullptr;       // front: estimate + demodulate
    cudaStream_t st_back = nullptr;  // back: tracker + decoder + log
    cudaStream_t st_copy = nullptr;  // host -> device sample copies
    cudaEvent_t ev_copy = nullptr;
    bool copy_pending = false;
    std::string cuda_err;

    // input
    uint32_t* d_iq_owned = nullptr;
    const uint32_t* d_iq = nullptr;
    int64_t stride = 0;
    bool ring = false, attached = false;
    std::vector<int64_t> h_avail;    // host truth: samples pushed / attached per stream
    int64_t* h_snap = nullptr;       // pinned + mapped [kRuns][S]: avail as seen by each run
    int64_t* h_snap_dev = nullptr;   // the same memory as the device sees it
    int64_t* d_avail = nullptr;      // [kRuns][S]

    // runs
    long long run_seq = 0;           // runs enqueued so far
    cudaEvent_t ev_front[kRuns]{}, ev_back[kRuns]{};
    cudaEvent_t ev_t[kRuns][5]{};    // front start, after estimate, after demod, after track, after decode
    bool slot_timed[kRuns]{};        // events of this slot not yet folded into acc_ms
    int64_t run_syms[kRuns]{};       // bound on the soft symbols each run can have produced
    double acc_ms[4]{};              // estimate, demod, track, decode since the last reset
    cudaEvent_t ev_first = nullptr, ev_last = nullptr;
    // OPVD_TRACE=1 (development aid): timeline of the pushes and runs since the last reset, printed by opvd_poll_frames
    bool trace = false;
    std::vector<cudaEvent_t> tr_c0, tr_c1;  // begin / end of every push on the copy stream
    int tr_n = 0;
    bool have_first = false, have_times = false;

    // per-stream state
    DemodState* d_dstate = nullptr;
    TrackState* d_tstate = nullptr;
    double* d_est = nullptr;
    int64_t* d_nsym = nullptr;       // [2][S] symbols available after a run, by run parity

    // soft symbols
    double* d_soft = nullptr;
    int64_t soft_stride = 0;
    bool soft_ring = false;

    // frames / events / tasks
    int max_frames = 0, max_events = 0, max_tasks = 0;
    FrameRec* d_frec = nullptr;
    uint8_t* d_frames = nullptr;
    int32_t* d_metrics = nullptr;
    TrackEvent* d_events = nullptr;
    int32_t* d_nevents = nullptr;
    FrameTask* d_tasks = nullptr;
    int32_t* d_ntasks = nullptr;
    unsigned long long* d_counters = nullptr;
    FrameLogEntry* d_log = nullptr;
    int64_t log_cap = 0;
    unsigned long long* d_log_count = nullptr;
    unsigned long long* h_log_count = nullptr;  // pinned mirror, refreshed after every run
    unsigned long long polled_log = 0, lost_frames = 0;
    std::vector<FrameLogEntry> fetch;           // staging of one poll
    std::deque<FrameLogEntry> pending;          // fetched, not yet handed to the caller
    std::vector<int32_t> polled_events;

    bool final_seen = false;
    bool est_all_done = false;  // every stream's offset estimate has been decided by an enqueued run: no more est launches
};

namespace {

#define CK(expr)                                                                  \
    do {                                                                          \
        cudaError_t _e = (expr);                                                  \
        if (_e != cudaSuccess) {                                                  \
            if (h) h->cuda_err = std::string(#expr) + ": " + cudaGetErrorString(_e); \
            return OPVD_ERR_CUDA;                                                 \
        }                                                                         \
    } while (0)

int use_device(int dev) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) return OPVD_ERR_CUDA;
    if (dev >= n) return OPVD_ERR_ARG;
    if (dev >= 0 && cudaSetDevice(dev) != cudaSuccess) return OPVD_ERR_CUDA;
    return OPVD_OK;
}

template <class T>
cudaError_t dalloc(T** p, size_t n) {
    return cudaMalloc(reinterpret_cast<void**>(p), std::max<size_t>(n, 1) * sizeof(T));
}

__global__ void init_state_kernel(DemodState* d, TrackState* t, double* est, int64_t* nsym, int n, int have_init,
                                  double init_off) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    DemodState s;
    demod_state_init(s);
    if (have_init) { s.freq_offset = init_off; s.flags |= kFlagEstDone; }
    d[i] = s;
    TrackState ts;
    track_state_init(ts);
    t[i] = ts;
    est[i] = 0.0;
    nsym[i] = 0;
    nsym[n + i] = 0;
}

// A run's per-stream sample counts go from the pinned snapshot to the device through a KERNEL that reads the host
// memory directly, not through cudaMemcpyAsync: on the copy engine an 8-byte-per-stream copy queues behind the next
// tile's 10 ms push (the engine arbitrates between streams, not by age), and the run that needs it starts one tile late
// (measured with OPVD_TRACE: the e2e leg lost 2-3 ms of PCIe time per tile to it).
__global__ void snapshot_kernel(int64_t* __restrict__ dst, const int64_t* __restrict__ src_host, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src_host[i];
}

bool stream_ring(const opvd_handle* h) { return h->cfg.mode == OPVD_MODE_STREAM && !h->attached; }

int ensure_output_buffers(opvd_handle* h, int64_t sample_capacity) {
    if (h->d_soft) return OPVD_OK;
    h->soft_ring = stream_ring(h);
    int64_t max_sym;
    if (h->cfg.max_symbols > 0) max_sym = h->cfg.max_symbols;
    else if (h->soft_ring)  // two tiles of at most one ring each in flight, plus what the tracker may reach back to
        max_sym = 2 * ((int64_t)((double)sample_capacity / kMinSamplesPerSymbol) + 2) + kTrackBack + 64;
    else
        max_sym = (int64_t)((double)sample_capacity / kMinSamplesPerSymbol) + 8;
    h->soft_stride = (max_sym + 3) & ~3ll;
    h->max_frames = h->cfg.max_frames > 0 ? h->cfg.max_frames : (int)(max_sym / kFrameSymbols + 2);
    h->max_events = 3 * h->max_frames + 64;
    const long long mt = (long long)h->S * h->max_frames;
    h->max_tasks = (int)std::min<long long>(mt, 1ll << 30);
    h->log_cap = h->max_tasks;
    CK(dalloc(&h->d_soft, (size_t)h->S * h->soft_stride));
    CK(dalloc(&h->d_frec, (size_t)h->S * h->max_frames));
    CK(dalloc(&h->d_frames, (size_t)h->S * h->max_frames * kFrameBytes));
    CK(dalloc(&h->d_metrics, (size_t)h->S * h->max_frames));
    CK(dalloc(&h->d_events, (size_t)h->S * h->max_events));
    CK(dalloc(&h->d_tasks, (size_t)h->max_tasks));
    CK(dalloc(&h->d_log, (size_t)h->log_cap));
    CK(cudaMemsetAsync(h->d_metrics, 0xFF, (size_t)h->S * h->max_frames * sizeof(int32_t), h->st));
    CK(cudaStreamSynchronize(h->st));
    return OPVD_OK;
}

// fold the kernel times of a finished run into the accumulators (its events are about to be reused)
int retire_slot(opvd_handle* h, int slot) {
    if (!h->slot_timed[slot]) return OPVD_OK;
    CK(cudaEventSynchronize(h->ev_back[slot]));
    float ms = 0.f;
    for (int i = 0; i < 4; ++i) {
        CK(cudaEventElapsedTime(&ms, h->ev_t[slot][i], h->ev_t[slot][i + 1]));
        h->acc_ms[i] += ms;
    }
    h->slot_timed[slot] = false;
    return OPVD_OK;
}

int wait_all(opvd_handle* h) {
    CK(cudaStreamSynchronize(h->st_copy));
    CK(cudaStreamSynchronize(h->st));
    CK(cudaStreamSynchronize(h->st_back));
    return OPVD_OK;
}

// first sample of stream s that must stay intact once run r (and all earlier runs) have demodulated
int64_t retained_start(const opvd_handle* h, long long r, int s) {
    if (r < 0 || h->cfg.mode != OPVD_MODE_STREAM) return 0;
    const int64_t a = h->h_snap[(size_t)(r % kRuns) * h->S + s];
    return std::max<int64_t>(0, a - kCarry);
}

}  // namespace

extern "C" {

int opvd_version(void) { return 200; }

const char* opvd_strerror(int code) {
    switch (code) {
        case OPVD_OK: return "ok";
        case OPVD_ERR_ARG: return "invalid argument";
        case OPVD_ERR_CUDA: return "CUDA error or no usable CUDA device (this library has no CPU fallback)";
        case OPVD_ERR_CAPACITY: return "buffer capacity exceeded";
        case OPVD_ERR_STATE: return "call not valid in the current state";
        case OPVD_ERR_ALIGN: return "device buffer must be 16-byte aligned with a stride multiple of 4 samples";
    }
    return "unknown error";
}

const char* opvd_last_cuda_error(const opvd_handle* h) { return h ? h->cuda_err.c_str() : ""; }

int opvd_create(const opvd_config* cfg, opvd_handle** out) {
    if (!cfg || !out || cfg->n_streams <= 0) return OPVD_ERR_ARG;
    if (cfg->mode != OPVD_MODE_BATCH && cfg->mode != OPVD_MODE_STREAM) return OPVD_ERR_ARG;
    int rc = use_device(cfg->device);
    if (rc != OPVD_OK) return rc;
    opvd_handle* h = new opvd_handle();
    h->cfg = *cfg;
    h->S = cfg->n_streams;
    cudaGetDevice(&h->dev);
    auto fail = [&](int code) { opvd_destroy(h); return code; };
    auto ok = [](cudaError_t e) { return e == cudaSuccess; };
    if (!ok(cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking)) ||
        !ok(cudaStreamCreateWithFlags(&h->st_back, cudaStreamNonBlocking)) ||
        !ok(cudaStreamCreateWithFlags(&h->st_copy, cudaStreamNonBlocking)) ||
        !ok(cudaEventCreateWithFlags(&h->ev_copy, cudaEventDisableTiming)) || !ok(cudaEventCreate(&h->ev_first)) ||
        !ok(cudaEventCreate(&h->ev_last)))
        return fail(OPVD_ERR_CUDA);
    h->trace = getenv("OPVD_TRACE") != nullptr;
    if (h->trace) {
        h->tr_c0.resize(64); h->tr_c1.resize(64);
        for (int i = 0; i < 64; ++i) { cudaEventCreate(&h->tr_c0[i]); cudaEventCreate(&h->tr_c1[i]); }
    }
    for (int r = 0; r < kRuns; ++r) {
        if (!ok(cudaEventCreateWithFlags(&h->ev_front[r], cudaEventDisableTiming)) ||
            !ok(cudaEventCreateWithFlags(&h->ev_back[r], cudaEventDisableTiming)))
            return fail(OPVD_ERR_CUDA);
        for (auto& e : h->ev_t[r])
            if (!ok(cudaEventCreate(&e))) return fail(OPVD_ERR_CUDA);
    }
    upload_constants();
    if (!ok(dalloc(&h->d_dstate, h->S)) || !ok(dalloc(&h->d_tstate, h->S)) || !ok(dalloc(&h->d_est, h->S)) ||
        !ok(dalloc(&h->d_avail, (size_t)kRuns * h->S)) || !ok(dalloc(&h->d_nsym, (size_t)2 * h->S)) ||
        !ok(dalloc(&h->d_nevents, h->S)) || !ok(dalloc(&h->d_ntasks, 1)) || !ok(dalloc(&h->d_counters, kNumCounters)) ||
        !ok(dalloc(&h->d_log_count, 1)) ||
        !ok(cudaHostAlloc(reinterpret_cast<void**>(&h->h_snap), sizeof(int64_t) * kRuns * h->S, cudaHostAllocMapped)) ||
        !ok(cudaHostGetDevicePointer(reinterpret_cast<void**>(&h->h_snap_dev), h->h_snap, 0)) ||
        !ok(cudaMallocHost(reinterpret_cast<void**>(&h->h_log_count), sizeof(unsigned long long))))
        return fail(OPVD_ERR_CUDA);
    *h->h_log_count = 0;
    cudaMemsetAsync(h->d_nevents, 0, sizeof(int32_t) * h->S, h->st);
    cudaMemsetAsync(h->d_ntasks, 0, sizeof(int32_t), h->st);
    cudaMemsetAsync(h->d_counters, 0, sizeof(unsigned long long) * kNumCounters, h->st);
    cudaMemsetAsync(h->d_log_count, 0, sizeof(unsigned long long), h->st);
    const int have_init = (cfg->mode == OPVD_MODE_STREAM && cfg->have_init_offset) ? 1 : 0;  // :1004 vs :1164
    init_state_kernel<<<(h->S + 127) / 128, 128, 0, h->st>>>(h->d_dstate, h->d_tstate, h->d_est, h->d_nsym, h->S,
                                                            have_init, cfg->init_offset_hz);
    h->h_avail.assign(h->S, 0);
    h->polled_events.assign(h->S, 0);
    h->est_all_done = have_init != 0;  // -o: the estimate is skipped altogether (:1031)
    if (cfg->max_samples > 0) {
        h->stride = (cfg->max_samples + 63) & ~63ll;
        h->ring = cfg->mode == OPVD_MODE_STREAM;
        if (h->ring && h->stride < kCarry + 64) h->stride = (kCarry + 64 + 63) & ~63ll;  // at least one chunk + carry
        if (!ok(dalloc(&h->d_iq_owned, (size_t)h->S * h->stride))) return fail(OPVD_ERR_CUDA);
        h->d_iq = h->d_iq_owned;
        if (ensure_output_buffers(h, h->stride) != OPVD_OK) return fail(OPVD_ERR_CUDA);
    }
    if (cudaStreamSynchronize(h->st) != cudaSuccess || cudaGetLastError() != cudaSuccess) return fail(OPVD_ERR_CUDA);
    *out = h;
    return OPVD_OK;
}

int opvd_destroy(opvd_handle* h) {
    if (!h) return OPVD_OK;
    cudaSetDevice(h->dev);
    if (h->st_copy) cudaStreamSynchronize(h->st_copy);
    if (h->st) cudaStreamSynchronize(h->st);
    if (h->st_back) cudaStreamSynchronize(h->st_back);
    cudaFree(h->d_iq_owned); cudaFree(h->d_avail); cudaFree(h->d_dstate); cudaFree(h->d_tstate); cudaFree(h->d_est);
    cudaFree(h->d_nsym); cudaFree(h->d_soft); cudaFree(h->d_frec); cudaFree(h->d_frames); cudaFree(h->d_metrics);
    cudaFree(h->d_events); cudaFree(h->d_nevents); cudaFree(h->d_tasks); cudaFree(h->d_ntasks); cudaFree(h->d_counters);
    cudaFree(h->d_log); cudaFree(h->d_log_count);
    if (h->h_snap) cudaFreeHost(h->h_snap);
    if (h->h_log_count) cudaFreeHost(h->h_log_count);
    for (int r = 0; r < kRuns; ++r) {
        if (h->ev_front[r]) cudaEventDestroy(h->ev_front[r]);
        if (h->ev_back[r]) cudaEventDestroy(h->ev_back[r]);
        for (auto& e : h->ev_t[r])
            if (e) cudaEventDestroy(e);
    }
    if (h->ev_first) cudaEventDestroy(h->ev_first);
    if (h->ev_last) cudaEventDestroy(h->ev_last);
    if (h->ev_copy) cudaEventDestroy(h->ev_copy);
    if (h->st_copy) cudaStreamDestroy(h->st_copy);
    if (h->st_back) cudaStreamDestroy(h->st_back);
    if (h->st) cudaStreamDestroy(h->st);
    delete h;
    return OPVD_OK;
}

int opvd_reset(opvd_handle* h) {
    if (!h) return OPVD_ERR_ARG;
    CK(cudaSetDevice(h->dev));
    int rc = wait_all(h);
    if (rc != OPVD_OK) return rc;
    h->copy_pending = false;
    const int have_init = (h->cfg.mode == OPVD_MODE_STREAM && h->cfg.have_init_offset) ? 1 : 0;
    init_state_kernel<<<(h->S + 127) / 128, 128, 0, h->st>>>(h->d_dstate, h->d_tstate, h->d_est, h->d_nsym, h->S,
                                                            have_init, h->cfg.init_offset_hz);
    CK(cudaMemsetAsync(h->d_nevents, 0, sizeof(int32_t) * h->S, h->st));
    CK(cudaMemsetAsync(h->d_ntasks, 0, sizeof(int32_t), h->st));
    CK(cudaMemsetAsync(h->d_counters, 0, sizeof(unsigned long long) * kNumCounters, h->st));
    CK(cudaMemsetAsync(h->d_log_count, 0, sizeof(unsigned long long), h->st));
    if (h->d_metrics)
        CK(cudaMemsetAsync(h->d_metrics, 0xFF, (size_t)h->S * h->max_frames * sizeof(int32_t), h->st));
    h->tr_n = 0;
    *h->h_log_count = 0;
    h->polled_log = 0;
    h->lost_frames = 0;
    h->pending.clear();
    std::fill(h->polled_events.begin(), h->polled_events.end(), 0);
    if (!h->attached) std::fill(h->h_avail.begin(), h->h_avail.end(), 0);  // attached captures stay attached
    h->run_seq = 0;
    for (int r = 0; r < kRuns; ++r) { h->slot_timed[r] = false; h->run_syms[r] = 0; }
    for (double& a : h->acc_ms) a = 0.0;
    h->have_first = h->have_times = false;
    h->final_seen = false;
    h->est_all_done = h->cfg.mode == OPVD_MODE_STREAM && h->cfg.have_init_offset;
    CK(cudaStreamSynchronize(h->st));
    CK(cudaGetLastError());
    return OPVD_OK;
}

static int push_common(opvd_handle* h, int32_t first, int32_t count, const int16_t* iq, int64_t n, int64_t host_stride) {
    if (!h || n < 0 || (n > 0 && !iq)) return OPVD_ERR_ARG;
    if (h->attached || !h->d_iq_owned) return OPVD_ERR_STATE;
    if (h->final_seen) return OPVD_ERR_STATE;
    if (n == 0) return OPVD_OK;
    CK(cudaSetDevice(h->dev));
    if (n > h->stride) return OPVD_ERR_CAPACITY;
    // ---- room: find the oldest run whose completion frees enough of the ring (linear buffers never free anything)
    auto fits = [&](long long r) {
        for (int s = first; s < first + count; ++s)
            if (h->h_avail[s] + n - (h->ring ? retained_start(h, r, s) : 0) > h->stride) return false;
        return true;
    };
    long long need = -2;  // run whose demodulator must have finished before the copy may start; -1: none
    for (int attempt = 0; attempt < 2 && need == -2; ++attempt) {
        if (h->run_seq == 0 || fits(-1)) {  // room without anything retiring: no wait
            if (fits(-1)) need = -1;
        } else {
            for (long long r = std::max<long long>(0, h->run_seq - kRuns); r < h->run_seq; ++r)
                if (fits(r)) { need = r; break; }
        }
        if (need == -2 && attempt == 0) {
            if (!h->ring) return OPVD_ERR_CAPACITY;
            // consume what is already there (an implicit run, like the reference's chunk loop), then look again
            bool unrun = h->run_seq == 0;
            for (int s = 0; s < h->S && !unrun; ++s)
                unrun = h->h_avail[s] > h->h_snap[(size_t)((h->run_seq - 1) % kRuns) * h->S + s];
            if (!unrun) return OPVD_ERR_CAPACITY;
            int rc = opvd_run(h, 0);
            if (rc != OPVD_OK) return rc;
        }
    }
    if (need == -2) return OPVD_ERR_CAPACITY;
    if (need >= 0) CK(cudaStreamWaitEvent(h->st_copy, h->ev_front[need % kRuns], 0));
    if (h->trace && h->tr_n < (int)h->tr_c0.size()) CK(cudaEventRecord(h->tr_c0[h->tr_n], h->st_copy));
    // ---- copy, in two pieces where the ring wraps
    bool uniform = true;
    for (int s = first; s < first + count; ++s) uniform = uniform && h->h_avail[s] == h->h_avail[first];
    auto copy_rows = [&](int s0, int rows, int64_t host_off, int64_t dst_off, int64_t len) -> cudaError_t {
        if (len <= 0) return cudaSuccess;
        if (rows == 1)
            return cudaMemcpyAsync(h->d_iq_owned + (size_t)s0 * h->stride + dst_off,
                                   iq + ((size_t)(s0 - first) * host_stride + host_off) * 2, (size_t)len * 4,
                                   cudaMemcpyHostToDevice, h->st_copy);
        return cudaMemcpy2DAsync(h->d_iq_owned + (size_t)s0 * h->stride + dst_off, (size_t)h->stride * 4,
                                 iq + ((size_t)(s0 - first) * host_stride + host_off) * 2, (size_t)host_stride * 4,
                                 (size_t)len * 4, (size_t)rows, cudaMemcpyHostToDevice, h->st_copy);
    };
    auto copy_span = [&](int s0, int rows) -> cudaError_t {
        const int64_t off = h->ring ? h->h_avail[s0] % h->stride : h->h_avail[s0];
        const int64_t len1 = std::min<int64_t>(n, h->stride - off);
        cudaError_t e = copy_rows(s0, rows, 0, off, len1);
        if (e != cudaSuccess) return e;
        return copy_rows(s0, rows, len1, 0, n - len1);
    };
    if (uniform) {
        CK(copy_span(first, count));
    } else {
        for (int s = first; s < first + count; ++s) CK(copy_span(s, 1));
    }
    // the next opvd_run waits for this event on the device before its kernels read the samples
    CK(cudaEventRecord(h->ev_copy, h->st_copy));
    if (h->trace && h->tr_n < (int)h->tr_c0.size()) CK(cudaEventRecord(h->tr_c1[h->tr_n++], h->st_copy));
    h->copy_pending = true;
    for (int s = first; s < first + count; ++s) h->h_avail[s] += n;
    return OPVD_OK;
}

int opvd_push_iq(opvd_handle* h, int32_t stream, const int16_t* iq, int64_t n_samples) {
    if (!h || stream < 0 || stream >= h->S) return OPVD_ERR_ARG;
    return push_common(h, stream, 1, iq, n_samples, n_samples);
}

int opvd_push_iq_all(opvd_handle* h, const int16_t* iq, int64_t n_samples, int64_t host_stride_samples) {
    if (!h || host_stride_samples < n_samples) return OPVD_ERR_ARG;
    return push_common(h, 0, h->S, iq, n_samples, host_stride_samples);
}

int opvd_attach_device_iq(opvd_handle* h, const void* d_iq, int64_t stride_samples, const int64_t* n_samples,
                          int64_t n_uniform) {
    if (!h || !d_iq || stride_samples <= 0) return OPVD_ERR_ARG;
    if (h->d_iq_owned) return OPVD_ERR_STATE;
    if ((reinterpret_cast<uintptr_t>(d_iq) & 15) || (stride_samples & 3)) return OPVD_ERR_ALIGN;
    if (h->attached && (h->d_iq != d_iq || h->stride != stride_samples)) return OPVD_ERR_STATE;
    CK(cudaSetDevice(h->dev));
    for (int s = 0; s < h->S; ++s) {
        const int64_t n = n_samples ? n_samples[s] : n_uniform;
        if (n < 0 || n > stride_samples) return OPVD_ERR_ARG;
        if (n < h->h_avail[s] && h->run_seq > 0) return OPVD_ERR_ARG;  // a stream never shrinks once it has been run
    }
    for (int s = 0; s < h->S; ++s) h->h_avail[s] = n_samples ? n_samples[s] : n_uniform;
    h->d_iq = static_cast<const uint32_t*>(d_iq);
    h->stride = stride_samples;
    h->ring = false;
    h->attached = true;
    return ensure_output_buffers(h, stride_samples);
}

int opvd_run(opvd_handle* h, int final_flag) {
    if (!h) return OPVD_ERR_ARG;
    if (!h->d_iq) return OPVD_ERR_STATE;
    CK(cudaSetDevice(h->dev));
    int rc = ensure_output_buffers(h, h->stride);
    if (rc != OPVD_OK) return rc;
    const long long t = h->run_seq;
    const int slot = (int)(t % kRuns);
    // ---- soft-symbol room for what this run can produce (bounded on the host: a symbol consumes >= 39.895 samples)
    int64_t new_syms = 0, max_avail = 0;
    for (int s = 0; s < h->S; ++s) {
        const int64_t from = retained_start(h, t - 1, s);
        new_syms = std::max(new_syms, (int64_t)((double)(h->h_avail[s] - from) / kMinSamplesPerSymbol) + 2);
        max_avail = std::max(max_avail, h->h_avail[s]);
    }
    bool serialize = false;  // front(t) must wait for back(t-1): the ring has room for one tile only
    if (!h->soft_ring) {
        if ((int64_t)((double)max_avail / kMinSamplesPerSymbol) + 8 > h->soft_stride) return OPVD_ERR_CAPACITY;
    } else {
        const int64_t prev = t > 0 ? h->run_syms[(t - 1) % kRuns] : 0;
        if (new_syms + prev + kTrackBack > h->soft_stride) {
            serialize = true;
            if (new_syms + kTrackBack > h->soft_stride) return OPVD_ERR_CAPACITY;
        }
    }
    // ---- this run's bookkeeping slot (its previous user finished long ago; fold its kernel times first)
    rc = retire_slot(h, slot);
    if (rc != OPVD_OK) return rc;
    if (t >= kRuns) CK(cudaEventSynchronize(h->ev_back[slot]));
    h->run_syms[slot] = new_syms;
    int64_t* snap = h->h_snap + (size_t)slot * h->S;
    std::copy(h->h_avail.begin(), h->h_avail.end(), snap);
    int64_t* d_avail = h->d_avail + (size_t)slot * h->S;
    prefer_max_shared(snapshot_kernel);
    snapshot_kernel<<<(h->S + 255) / 256, 256, 0, h->st>>>(d_avail, h->h_snap_dev + (size_t)slot * h->S, h->S);
    if (h->copy_pending) {  // pushed samples must have landed before the kernels read them (device-side wait)
        CK(cudaStreamWaitEvent(h->st, h->ev_copy, 0));
        h->copy_pending = false;
    }
    // demod(t) reuses the n_sym array of run t-2 and, in a ring, soft rows that back(t-2) may still be reading
    if (t >= 2) CK(cudaStreamWaitEvent(h->st, h->ev_back[(t - 2) % kRuns], 0));
    if (serialize && t >= 1) CK(cudaStreamWaitEvent(h->st, h->ev_back[(t - 1) % kRuns], 0));

    StreamBuffers sb{h->d_iq, h->stride, d_avail, h->ring ? 1 : 0};
    SoftBuffers so{h->d_soft, h->soft_stride, h->soft_ring ? 1 : 0, h->d_nsym + (size_t)(t & 1) * h->S};
    // ---- front: estimate + demodulate
    if (!h->have_first) {
        CK(cudaEventRecord(h->ev_first, h->st));
        h->have_first = true;
    }
    CK(cudaEventRecord(h->ev_t[slot][0], h->st));
    if (!h->est_all_done) {
        launch_estimate(sb, h->d_dstate, h->d_est, h->S, h->cfg.mode, final_flag ? 1 : 0, h->st);
        // the estimate of a stream is decided by the first run that sees a full chunk of it (stream mode, :1030-1038)
        // or by the final run (batch mode :1166; short streams are never estimated): once that holds for every stream
        // the kernel has nothing left to do
        int64_t min_avail = INT64_MAX;
        for (int s = 0; s < h->S; ++s) min_avail = std::min(min_avail, h->h_avail[s]);
        h->est_all_done = final_flag || (h->cfg.mode == OPVD_MODE_STREAM && min_avail >= kChunkSamples);
    }
    CK(cudaEventRecord(h->ev_t[slot][1], h->st));
    if (h->cfg.coherent && h->cfg.mode == OPVD_MODE_BATCH) {
        // -p <hz> verbatim like the reference (set_pll_bandwidth, :1149; -p 0 freezes the loop); NaN = default 50 (:946)
        const double bw = std::isnan(h->cfg.pll_bw_hz) ? 50.0 : h->cfg.pll_bw_hz;
        CK(launch_demod_coherent(sb, so, h->d_dstate, h->S, final_flag ? 1 : 0, h->cfg.afc_alpha, bw, h->d_counters, h->st));
    } else {
        CK(launch_demod(sb, so, h->d_dstate, h->S, h->cfg.mode, final_flag ? 1 : 0, h->cfg.afc_alpha,
                        h->cfg.lanes_per_stream, h->d_counters, h->st));
    }
    CK(cudaEventRecord(h->ev_t[slot][2], h->st));
    CK(cudaEventRecord(h->ev_front[slot], h->st));
    // ---- back: tracker + decoder + frame log, overlapping the front of the next run
    CK(cudaStreamWaitEvent(h->st_back, h->ev_front[slot], 0));
    CK(cudaMemsetAsync(h->d_ntasks, 0, sizeof(int32_t), h->st_back));
    launch_track(so, h->d_tstate, h->S, h->d_frec, h->max_frames, h->d_events, h->d_nevents, h->max_events, h->d_tasks,
                 h->d_ntasks, h->max_tasks, h->d_counters, h->st_back);
    CK(cudaEventRecord(h->ev_t[slot][3], h->st_back));
    launch_decode(so, h->d_tasks, h->d_ntasks, h->max_tasks, h->d_frames, h->d_metrics, h->max_frames, h->d_frec, h->d_log,
                  h->d_log_count, h->log_cap, h->d_counters, h->st_back);
    CK(cudaEventRecord(h->ev_t[slot][4], h->st_back));
    CK(cudaMemcpyAsync(h->h_log_count, h->d_log_count, sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->st_back));
    CK(cudaEventRecord(h->ev_back[slot], h->st_back));
    CK(cudaEventRecord(h->ev_last, h->st_back));
    CK(cudaGetLastError());
    h->slot_timed[slot] = true;
    h->have_times = true;
    h->run_seq = t + 1;
    if (final_flag) h->final_seen = true;
    return OPVD_OK;
}

int opvd_sync(opvd_handle* h) {
    if (!h) return OPVD_ERR_ARG;
    CK(cudaSetDevice(h->dev));
    int rc = wait_all(h);
    if (rc != OPVD_OK) return rc;
    CK(cudaGetLastError());
    return OPVD_OK;
}

int opvd_demod_lanes(opvd_handle* h) {
    if (!h) return OPVD_ERR_ARG;
    CK(cudaSetDevice(h->dev));
    return h->cfg.lanes_per_stream > 0 ? h->cfg.lanes_per_stream : demod_auto_lanes(h->S);
}

int opvd_last_run_ms(opvd_handle* h, float* ms5) {
    if (!h || !ms5) return OPVD_ERR_ARG;
    if (!h->have_times) return OPVD_ERR_STATE;
    CK(cudaSetDevice(h->dev));
    for (int r = 0; r < kRuns; ++r) {
        int rc = retire_slot(h, r);
        if (rc != OPVD_OK) return rc;
    }
    for (int i = 0; i < 4; ++i) ms5[i] = (float)h->acc_ms[i];
    CK(cudaEventSynchronize(h->ev_last));
    CK(cudaEventElapsedTime(&ms5[4], h->ev_first, h->ev_last));
    return OPVD_OK;
}

static int poll_frames_impl(opvd_handle* h, int32_t max_frames, uint8_t* frames134, opvd_frame_info* info, bool wait) {
    if (!h || max_frames < 0 || (max_frames > 0 && !frames134)) return OPVD_ERR_ARG;
    CK(cudaSetDevice(h->dev));
    if (!h->d_log) return 0;
    // ---- fetch what the runs enqueued so far have logged since the last poll
    if (wait && h->run_seq > 0) CK(cudaEventSynchronize(h->ev_back[(h->run_seq - 1) % kRuns]));
    if (wait && h->trace && h->run_seq > 0) {
        CK(cudaStreamSynchronize(h->st_copy));
        float a = 0.f, b = 0.f;
        const cudaEvent_t base = h->tr_n > 0 ? h->tr_c0[0] : h->ev_t[0][0];
        for (int i = 0; i < h->tr_n; ++i) {
            cudaEventElapsedTime(&a, base, h->tr_c0[i]);
            cudaEventElapsedTime(&b, base, h->tr_c1[i]);
            fprintf(stderr, "opvd trace: push %d copy %.2f .. %.2f ms\n", i, a, b);
        }
        for (long long r = 0; r < h->run_seq && h->run_seq <= kRuns; ++r) {  // the run events live in kRuns slots
            float t[5];
            for (int i = 0; i < 5; ++i) cudaEventElapsedTime(&t[i], base, h->ev_t[r % kRuns][i]);
            fprintf(stderr, "opvd trace: run %lld front %.2f est %.2f demod %.2f | back track %.2f decode %.2f ms\n", r, t[0], t[1], t[2], t[3], t[4]);
        }
        h->tr_n = 0;
    }
    const unsigned long long total = *h->h_log_count;
    if (total - h->polled_log > (unsigned long long)h->log_cap) {  // the log wrapped over frames nobody polled
        const unsigned long long lost = total - h->polled_log - (unsigned long long)h->log_cap;
        h->lost_frames += lost;
        h->polled_log += lost;
    }
    const size_t n_new = (size_t)(total - h->polled_log);
    if (n_new) {
        h->fetch.resize(n_new);
        const size_t p0 = (size_t)(h->polled_log % (unsigned long long)h->log_cap);
        const size_t n1 = std::min(n_new, (size_t)h->log_cap - p0);
        CK(cudaMemcpy(h->fetch.data(), h->d_log + p0, n1 * sizeof(FrameLogEntry), cudaMemcpyDeviceToHost));
        if (n_new > n1)
            CK(cudaMemcpy(h->fetch.data() + n1, h->d_log, (n_new - n1) * sizeof(FrameLogEntry), cudaMemcpyDeviceToHost));
        // hand them out by (stream, frame): sort 8-byte keys, not the 176-byte entries
        std::vector<std::pair<uint64_t, uint32_t>> order(n_new);
        for (size_t i = 0; i < n_new; ++i)
            order[i] = {((uint64_t)(uint32_t)h->fetch[i].stream << 32) | (uint32_t)h->fetch[i].frame_idx, (uint32_t)i};
        std::sort(order.begin(), order.end());
        for (const auto& o : order) {
            const FrameLogEntry& e = h->fetch[o.second];
            if (e.metric >= 0) h->pending.push_back(e);  // dropped frames (:1052) are never written by the reference
        }
        h->polled_log = total;
    }
    int n = 0;
    while (n < max_frames && !h->pending.empty()) {
        const FrameLogEntry& e = h->pending.front();
        memcpy(frames134 + (size_t)n * kFrameBytes, e.frame, kFrameBytes);
        if (info) {
            opvd_frame_info& fi = info[n];
            fi.stream = e.stream; fi.frame_idx = e.frame_idx; fi.metric = e.metric; fi.reserved = 0;
            fi.payload_start = e.payload_start; fi.ready_idx = e.ready_idx; fi.sync_quality = e.quality;
        }
        h->pending.pop_front();
        ++n;
    }
    return n;
}

int opvd_poll_frames(opvd_handle* h, int32_t max_frames, uint8_t* frames134, opvd_frame_info* info) {
    return poll_frames_impl(h, max_frames, frames134, info, true);
}

// the frames of the runs that have FINISHED: the log counter reaches the pinned host word at the end of every run's
// decoder, so nothing has to be waited for and the pushes that are queued keep the copy engine busy meanwhile
int opvd_poll_frames_ready(opvd_handle* h, int32_t max_frames, uint8_t* frames134, opvd_frame_info* info) {
    return poll_frames_impl(h, max_frames, frames134, info, false);
}

int opvd_frames_lost(opvd_handle* h, uint64_t* out) {
    if (!h || !out) return OPVD_ERR_ARG;
    *out = h->lost_frames;
    return OPVD_OK;
}

int opvd_poll_events(opvd_handle* h, int32_t stream, int32_t max_events, opvd_event* out) {
    if (!h || stream < 0 || stream >= h->S || max_events < 0 || (max_events > 0 && !out)) return OPVD_ERR_ARG;
    if (!h->d_events) return 0;
    CK(cudaSetDevice(h->dev));
    if (h->run_seq > 0) CK(cudaEventSynchronize(h->ev_back[(h->run_seq - 1) % kRuns]));
    int32_t total = 0;
    CK(cudaMemcpy(&total, h->d_nevents + stream, sizeof(int32_t), cudaMemcpyDeviceToHost));
    int32_t& polled = h->polled_events[stream];
    if (total - polled > h->max_events) polled = total - h->max_events;  // ring overrun: the oldest events are gone
    const int n = std::min<int>(max_events, total - polled);
    if (n <= 0) return 0;
    std::vector<TrackEvent> ev(n);
    const int p0 = polled % h->max_events, n1 = std::min(n, h->max_events - p0);
    const TrackEvent* row = h->d_events + (size_t)stream * h->max_events;
    CK(cudaMemcpy(ev.data(), row + p0, (size_t)n1 * sizeof(TrackEvent), cudaMemcpyDeviceToHost));
    if (n > n1) CK(cudaMemcpy(ev.data() + n1, row, (size_t)(n - n1) * sizeof(TrackEvent), cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; ++i) {
        out[i].type = ev[i].type; out[i].count = ev[i].count; out[i].sym_idx = ev[i].sym_idx;
        out[i].corr = ev[i].corr; out[i].raw = ev[i].raw;
    }
    polled += n;
    return n;
}

int opvd_get_soft(opvd_handle* h, int32_t stream, int64_t first_sym, int64_t n, double* out) {
    if (!h || stream < 0 || stream >= h->S || first_sym < 0 || n < 0 || (n > 0 && !out)) return OPVD_ERR_ARG;
    if (!h->d_soft) return 0;
    CK(cudaSetDevice(h->dev));
    CK(cudaStreamSynchronize(h->st));
    DemodState ds;
    CK(cudaMemcpy(&ds, h->d_dstate + stream, sizeof(ds), cudaMemcpyDeviceToHost));
    if (h->soft_ring && first_sym < ds.n_sym - h->soft_stride) return OPVD_ERR_ARG;  // already overwritten
    const int64_t m = std::max<int64_t>(0, std::min<int64_t>(n, ds.n_sym - first_sym));
    if (m > 0) {
        const double* row = h->d_soft + (size_t)stream * h->soft_stride;
        const int64_t p0 = h->soft_ring ? first_sym % h->soft_stride : first_sym;
        const int64_t m1 = std::min<int64_t>(m, h->soft_stride - p0);
        CK(cudaMemcpy(out, row + p0, (size_t)m1 * sizeof(double), cudaMemcpyDeviceToHost));
        if (m > m1) CK(cudaMemcpy(out + m1, row, (size_t)(m - m1) * sizeof(double), cudaMemcpyDeviceToHost));
    }
    return (int)std::min<int64_t>(m, INT32_MAX);
}

int opvd_get_stream_info(opvd_handle* h, int32_t stream, opvd_stream_info* out) {
    if (!h || !out || stream < 0 || stream >= h->S) return OPVD_ERR_ARG;
    CK(cudaSetDevice(h->dev));
    CK(cudaStreamSynchronize(h->st));
    CK(cudaStreamSynchronize(h->st_back));
    DemodState ds;
    TrackState ts;
    double est;
    CK(cudaMemcpy(&ds, h->d_dstate + stream, sizeof(ds), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&ts, h->d_tstate + stream, sizeof(ts), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&est, h->d_est + stream, sizeof(est), cudaMemcpyDeviceToHost));
    out->est_offset_hz = est; out->freq_offset_hz = ds.freq_offset; out->timing_freq = ds.timing_freq;
    out->n_symbols = ds.n_sym; out->n_samples_used = ds.origin; out->sync_state = ts.state;
    out->frames_ready = ts.total_frames; out->done = (ds.flags & kFlagDone) ? 1 : 0; out->reserved = 0;
    return OPVD_OK;
}

int opvd_get_counters(opvd_handle* h, uint64_t* out, int32_t n) {
    if (!h || !out || n <= 0) return OPVD_ERR_ARG;
    CK(cudaSetDevice(h->dev));
    CK(cudaStreamSynchronize(h->st));
    CK(cudaStreamSynchronize(h->st_back));
    unsigned long long tmp[kNumCounters];
    CK(cudaMemcpy(tmp, h->d_counters, sizeof(tmp), cudaMemcpyDeviceToHost));
    for (int i = 0; i < n && i < kNumCounters; ++i) out[i] = tmp[i];
    return OPVD_OK;
}

int opvd_counters_device_ptr(opvd_handle* h, void** out) {
    if (!h || !out) return OPVD_ERR_ARG;
    *out = h->d_counters;
    return OPVD_OK;
}

int opvd_stage_decode_dev(int32_t device, const double* d_payloads, int32_t n, uint8_t* d_frames134, int32_t* d_metrics,
                          float* ms) {
    opvd_handle* h = nullptr;
    if (n < 0 || (n > 0 && (!d_payloads || !d_frames134 || !d_metrics))) return OPVD_ERR_ARG;
    int rc = use_device(device);
    if (rc != OPVD_OK) return rc;
    upload_constants();
    unsigned long long* ctr = nullptr;
    CK(dalloc(&ctr, kNumCounters));
    CK(cudaMemset(ctr, 0, sizeof(unsigned long long) * kNumCounters));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0, 0));
    launch_decode_payloads(d_payloads, n, d_frames134, d_metrics, ctr, 0);
    CK(cudaEventRecord(e1, 0));
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    if (ms) CK(cudaEventElapsedTime(ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(ctr);
    return OPVD_OK;
}

int opvd_stage_decode(int32_t device, const double* payloads, int32_t n, uint8_t* frames134, int32_t* metrics) {
    opvd_handle* h = nullptr;
    if (n < 0 || (n > 0 && (!payloads || !frames134 || !metrics))) return OPVD_ERR_ARG;
    int rc = use_device(device);
    if (rc != OPVD_OK) return rc;
    if (n == 0) return OPVD_OK;
    double* dp = nullptr;
    uint8_t* df = nullptr;
    int32_t* dm = nullptr;
    CK(dalloc(&dp, (size_t)n * kEncodedBits));
    CK(dalloc(&df, (size_t)n * kFrameBytes));
    CK(dalloc(&dm, (size_t)n));
    CK(cudaMemcpy(dp, payloads, (size_t)n * kEncodedBits * sizeof(double), cudaMemcpyHostToDevice));
    rc = opvd_stage_decode_dev(device, dp, n, df, dm, nullptr);
    if (rc == OPVD_OK) {
        CK(cudaMemcpy(frames134, df, (size_t)n * kFrameBytes, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(metrics, dm, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost));
    }
    cudaFree(dp); cudaFree(df); cudaFree(dm);
    return rc;
}

static SynthParams to_params(const opvd_synth* p) {
    SynthParams sp{};
    sp.n_streams = p->n_streams; sp.n_frames = p->n_frames; sp.stride = p->stride_samples; sp.n_samples = p->n_samples;
    sp.seed = p->seed; sp.scale = p->scale; sp.ebn0_lo_db = p->ebn0_lo_db; sp.ebn0_hi_db = p->ebn0_hi_db;
    sp.cfo_max_hz = p->cfo_max_hz; sp.frac_delay = p->frac_delay; sp.max_lead = p->max_lead;
    sp.first_stream = p->first_stream;
    return sp;
}

int opvd_synth_bank(int32_t device, const opvd_synth* p, void* d_iq) {
    opvd_handle* h = nullptr;
    if (!p || !d_iq || p->n_streams <= 0 || p->n_frames < 0 || p->n_samples > p->stride_samples) return OPVD_ERR_ARG;
    int rc = use_device(device);
    if (rc != OPVD_OK) return rc;
    SynthParams sp = to_params(p);
    uint8_t* scratch = nullptr;
    CK(dalloc(&scratch, synth_scratch_bytes(sp)));
    uint8_t* syms = scratch;
    int8_t* fsign = reinterpret_cast<int8_t*>(scratch + (size_t)sp.n_streams * sp.n_frames * kFrameSymbols);
    launch_synth(sp, static_cast<uint32_t*>(d_iq), syms, fsign, 0);
    CK(cudaDeviceSynchronize());
    CK(cudaGetLastError());
    cudaFree(scratch);
    return OPVD_OK;
}

int opvd_bert_check(opvd_handle* h, const opvd_synth* p) {
    if (!h || !p) return OPVD_ERR_ARG;
    if (!h->d_frames) return OPVD_ERR_STATE;
    CK(cudaSetDevice(h->dev));
    CK(cudaStreamSynchronize(h->st));
    CK(cudaStreamSynchronize(h->st_back));
    std::vector<TrackState> ts(h->S);
    CK(cudaMemcpy(ts.data(), h->d_tstate, sizeof(TrackState) * h->S, cudaMemcpyDeviceToHost));
    std::vector<int32_t> nf(h->S);
    for (int s = 0; s < h->S; ++s) nf[s] = std::min(ts[s].total_frames, h->max_frames);
    int32_t* d_nf = nullptr;
    CK(dalloc(&d_nf, h->S));
    CK(cudaMemcpy(d_nf, nf.data(), sizeof(int32_t) * h->S, cudaMemcpyHostToDevice));
    launch_bert_check_impl(h->d_frames, h->d_metrics, h->d_frec, d_nf, h->S, h->max_frames, to_params(p),
                           h->d_counters, h->st_back);
    CK(cudaStreamSynchronize(h->st_back));
    CK(cudaGetLastError());
    cudaFree(d_nf);
    return OPVD_OK;
}

}  // extern "C"
