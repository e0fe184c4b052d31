// opvd_api.cu — host runtime behind include/opvd.h: owns the device buffers and the per-stream
// state of the receive chain, sequences the sm_100a kernels on one CUDA stream, and moves frames,
// events and counters back to caller-owned host memory.  No compute happens on the host and there
// is no CPU fallback: without a CUDA device every entry point returns OPVD_ERR_CUDA.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/opvd.h"
#include "opvd_kernels.cuh"

using namespace opvd;

static_assert(sizeof(opvd_event) == sizeof(TrackEvent), "event layout");
static_assert(OPVD_NUM_COUNTERS == kNumCounters, "counter layout");

struct opvd_handle {
    opvd_config cfg{};
    int dev = 0;
    int S = 0;
    cudaStream_t st = nullptr;
    // host->device copies of opvd_push_iq* run on their own stream, so that the samples of the next time tile cross
    // PCIe while the kernels of the previous opvd_run are still working; opvd_run waits for them on the device
    cudaStream_t st_copy = nullptr;
    cudaEvent_t ev_copy = nullptr;
    bool copy_pending = false;
    std::string cuda_err;

    // input
    uint32_t* d_iq_owned = nullptr;
    const uint32_t* d_iq = nullptr;
    int64_t stride = 0, row_base = 0;
    bool attached = false;
    std::vector<int64_t> h_avail;
    int64_t* d_avail = nullptr;
    bool avail_dirty = true;

    // per-stream state
    DemodState* d_dstate = nullptr;
    TrackState* d_tstate = nullptr;
    double* d_est = nullptr;

    // soft symbols
    double* d_soft = nullptr;
    int64_t soft_stride = 0, soft_base = 0;

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

    // host mirrors for polling
    std::vector<TrackState> h_tstate;
    std::vector<FrameRec> h_frec;
    std::vector<uint8_t> h_frames;
    std::vector<int32_t> h_metrics;
    std::vector<TrackEvent> h_events;
    std::vector<int32_t> h_nevents;
    std::vector<int32_t> polled_frames, polled_events;
    bool mirror_stale = true, ev_mirror_stale = true;

    cudaEvent_t ev[5]{};
    bool have_times = false;
    bool final_seen = false;
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

__global__ void init_state_kernel(DemodState* d, TrackState* t, double* est, int n, int have_init, double init_off) {
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
}

// shift every row left by `shift` elements, keeping `keep` elements (via scratch, rows may overlap themselves)
template <class T>
__global__ void row_copy_kernel(const T* __restrict__ src, T* __restrict__ dst, long long src_stride,
                                long long dst_stride, long long src_off, long long keep) {
    const long long row = blockIdx.y;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < keep; i += (long long)gridDim.x * blockDim.x)
        dst[row * dst_stride + i] = src[row * src_stride + src_off + i];
}

int ensure_output_buffers(opvd_handle* h, int64_t sample_capacity) {
    if (h->d_soft) return OPVD_OK;
    int64_t max_sym = h->cfg.max_symbols > 0 ? h->cfg.max_symbols : sample_capacity / kSps + 64;
    h->soft_stride = (max_sym + 3) & ~3ll;
    h->max_frames = h->cfg.max_frames > 0 ? h->cfg.max_frames : (int)(max_sym / kFrameSymbols + 2);
    h->max_events = 3 * h->max_frames + 64;
    const long long mt = (long long)h->S * h->max_frames;
    h->max_tasks = (int)std::min<long long>(mt, 1ll << 30);
    CK(dalloc(&h->d_soft, (size_t)h->S * h->soft_stride));
    CK(dalloc(&h->d_frec, (size_t)h->S * h->max_frames));
    CK(dalloc(&h->d_frames, (size_t)h->S * h->max_frames * kFrameBytes));
    CK(dalloc(&h->d_metrics, (size_t)h->S * h->max_frames));
    CK(dalloc(&h->d_events, (size_t)h->S * h->max_events));
    CK(dalloc(&h->d_tasks, (size_t)h->max_tasks));
    CK(cudaMemsetAsync(h->d_metrics, 0xFF, (size_t)h->S * h->max_frames * sizeof(int32_t), h->st));
    return OPVD_OK;
}

// stream-mode housekeeping for unbounded input: drop consumed samples / soft symbols from the front
int compact(opvd_handle* h) {
    CK(cudaStreamSynchronize(h->st_copy));  // rows are about to move: no push may still be writing into them
    CK(cudaStreamSynchronize(h->st));
    std::vector<DemodState> ds(h->S);
    CK(cudaMemcpy(ds.data(), h->d_dstate, sizeof(DemodState) * h->S, cudaMemcpyDeviceToHost));
    h->h_tstate.resize(h->S);
    CK(cudaMemcpy(h->h_tstate.data(), h->d_tstate, sizeof(TrackState) * h->S, cudaMemcpyDeviceToHost));
    int64_t min_origin = INT64_MAX, max_avail = 0, min_need = INT64_MAX, max_sym = 0;
    for (int s = 0; s < h->S; ++s) {
        min_origin = std::min(min_origin, ds[s].origin);
        max_avail = std::max(max_avail, h->h_avail[s]);
        const TrackState& t = h->h_tstate[s];
        int64_t need;
        if (t.state == kHunting) need = t.cursor - kSyncBits;
        else if (t.state == kVerifying || t.collecting) need = t.payload_start - 1;
        else need = t.anchor + kFrameSymbols - kSyncBits;
        min_need = std::min(min_need, std::max<int64_t>(need, 0));
        max_sym = std::max(max_sym, ds[s].n_sym);
    }
    // samples: keep everything from 64 samples before the earliest call origin
    if (!h->attached && h->d_iq_owned) {
        int64_t new_base = std::max<int64_t>(h->row_base, ((min_origin - 64) / 64) * 64);
        if (min_origin < 64) new_base = h->row_base;
        const int64_t shift = new_base - h->row_base;
        const int64_t keep = max_avail - new_base;
        if (shift > 0 && keep >= 0) {
            uint32_t* tmp = nullptr;
            CK(dalloc(&tmp, (size_t)h->S * std::max<int64_t>(keep, 1)));
            dim3 g((unsigned)std::min<int64_t>((keep + 255) / 256 + 1, 1024), (unsigned)h->S);
            row_copy_kernel<uint32_t><<<g, 256, 0, h->st>>>(h->d_iq_owned, tmp, h->stride, keep, shift, keep);
            row_copy_kernel<uint32_t><<<g, 256, 0, h->st>>>(tmp, h->d_iq_owned, keep, h->stride, 0, keep);
            CK(cudaStreamSynchronize(h->st));
            cudaFree(tmp);
            h->row_base = new_base;
        }
    }
    if (h->d_soft) {
        const int64_t new_sbase = std::max<int64_t>(h->soft_base, min_need & ~3ll);
        const int64_t shift = new_sbase - h->soft_base;
        const int64_t keep = max_sym - new_sbase;
        if (shift > 0 && keep >= 0) {
            double* tmp = nullptr;
            CK(dalloc(&tmp, (size_t)h->S * std::max<int64_t>(keep, 1)));
            dim3 g((unsigned)std::min<int64_t>((keep + 255) / 256 + 1, 1024), (unsigned)h->S);
            row_copy_kernel<double><<<g, 256, 0, h->st>>>(h->d_soft, tmp, h->soft_stride, keep, shift, keep);
            row_copy_kernel<double><<<g, 256, 0, h->st>>>(tmp, h->d_soft, keep, h->soft_stride, 0, keep);
            CK(cudaStreamSynchronize(h->st));
            cudaFree(tmp);
            h->soft_base = new_sbase;
        }
    }
    return OPVD_OK;
}

int refresh_frame_mirror(opvd_handle* h) {
    if (!h->mirror_stale) return OPVD_OK;
    CK(cudaStreamSynchronize(h->st));
    h->h_tstate.resize(h->S);
    CK(cudaMemcpy(h->h_tstate.data(), h->d_tstate, sizeof(TrackState) * h->S, cudaMemcpyDeviceToHost));
    if (h->d_frames) {
        const size_t n = (size_t)h->S * h->max_frames;
        h->h_frec.resize(n);
        h->h_metrics.resize(n);
        h->h_frames.resize(n * kFrameBytes);
        CK(cudaMemcpy(h->h_frec.data(), h->d_frec, n * sizeof(FrameRec), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(h->h_metrics.data(), h->d_metrics, n * sizeof(int32_t), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(h->h_frames.data(), h->d_frames, n * kFrameBytes, cudaMemcpyDeviceToHost));
    }
    h->mirror_stale = false;
    return OPVD_OK;
}

}  // namespace

extern "C" {

int opvd_version(void) { return 100; }

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
    if (cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking) != cudaSuccess) return fail(OPVD_ERR_CUDA);
    if (cudaStreamCreateWithFlags(&h->st_copy, cudaStreamNonBlocking) != cudaSuccess) return fail(OPVD_ERR_CUDA);
    if (cudaEventCreateWithFlags(&h->ev_copy, cudaEventDisableTiming) != cudaSuccess) return fail(OPVD_ERR_CUDA);
    for (auto& e : h->ev)
        if (cudaEventCreate(&e) != cudaSuccess) return fail(OPVD_ERR_CUDA);
    upload_constants();
    if (dalloc(&h->d_dstate, h->S) != cudaSuccess || dalloc(&h->d_tstate, h->S) != cudaSuccess ||
        dalloc(&h->d_est, h->S) != cudaSuccess || dalloc(&h->d_avail, h->S) != cudaSuccess ||
        dalloc(&h->d_nevents, h->S) != cudaSuccess || dalloc(&h->d_ntasks, 1) != cudaSuccess ||
        dalloc(&h->d_counters, kNumCounters) != cudaSuccess)
        return fail(OPVD_ERR_CUDA);
    cudaMemsetAsync(h->d_nevents, 0, sizeof(int32_t) * h->S, h->st);
    cudaMemsetAsync(h->d_ntasks, 0, sizeof(int32_t), h->st);
    cudaMemsetAsync(h->d_counters, 0, sizeof(unsigned long long) * kNumCounters, h->st);
    const int have_init = (cfg->mode == OPVD_MODE_STREAM && cfg->have_init_offset) ? 1 : 0;  // :1004 vs :1164
    init_state_kernel<<<(h->S + 127) / 128, 128, 0, h->st>>>(h->d_dstate, h->d_tstate, h->d_est, h->S, have_init,
                                                            cfg->init_offset_hz);
    h->h_avail.assign(h->S, 0);
    h->polled_frames.assign(h->S, 0);
    h->polled_events.assign(h->S, 0);
    if (cfg->max_samples > 0) {
        h->stride = (cfg->max_samples + 63) & ~63ll;
        if (dalloc(&h->d_iq_owned, (size_t)h->S * h->stride) != cudaSuccess) return fail(OPVD_ERR_CUDA);
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
    cudaFree(h->d_iq_owned); cudaFree(h->d_avail); cudaFree(h->d_dstate); cudaFree(h->d_tstate); cudaFree(h->d_est);
    cudaFree(h->d_soft); cudaFree(h->d_frec); cudaFree(h->d_frames); cudaFree(h->d_metrics); cudaFree(h->d_events);
    cudaFree(h->d_nevents); cudaFree(h->d_tasks); cudaFree(h->d_ntasks); cudaFree(h->d_counters);
    for (auto& e : h->ev)
        if (e) cudaEventDestroy(e);
    if (h->ev_copy) cudaEventDestroy(h->ev_copy);
    if (h->st_copy) cudaStreamDestroy(h->st_copy);
    if (h->st) cudaStreamDestroy(h->st);
    delete h;
    return OPVD_OK;
}

int opvd_reset(opvd_handle* h) {
    if (!h) return OPVD_ERR_ARG;
    CK(cudaSetDevice(h->dev));
    CK(cudaStreamSynchronize(h->st_copy));
    CK(cudaStreamSynchronize(h->st));
    h->copy_pending = false;
    const int have_init = (h->cfg.mode == OPVD_MODE_STREAM && h->cfg.have_init_offset) ? 1 : 0;
    init_state_kernel<<<(h->S + 127) / 128, 128, 0, h->st>>>(h->d_dstate, h->d_tstate, h->d_est, h->S, have_init,
                                                            h->cfg.init_offset_hz);
    CK(cudaMemsetAsync(h->d_nevents, 0, sizeof(int32_t) * h->S, h->st));
    CK(cudaMemsetAsync(h->d_ntasks, 0, sizeof(int32_t), h->st));
    CK(cudaMemsetAsync(h->d_counters, 0, sizeof(unsigned long long) * kNumCounters, h->st));
    if (h->d_metrics)
        CK(cudaMemsetAsync(h->d_metrics, 0xFF, (size_t)h->S * h->max_frames * sizeof(int32_t), h->st));
    std::fill(h->polled_frames.begin(), h->polled_frames.end(), 0);
    std::fill(h->polled_events.begin(), h->polled_events.end(), 0);
    if (!h->attached) {  // library-owned input starts empty again; attached captures stay attached
        std::fill(h->h_avail.begin(), h->h_avail.end(), 0);
        h->avail_dirty = true;
    }
    h->row_base = 0;
    h->soft_base = 0;
    h->mirror_stale = h->ev_mirror_stale = true;
    h->final_seen = false;
    h->have_times = false;
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
    int64_t max_avail = 0;
    for (int s = first; s < first + count; ++s) max_avail = std::max(max_avail, h->h_avail[s]);
    if (max_avail + n - h->row_base > h->stride) {
        if (h->cfg.mode == OPVD_MODE_STREAM) {
            int rc = opvd_run(h, 0);  // consume what is already there, then drop it from the front
            if (rc != OPVD_OK) return rc;
            rc = compact(h);
            if (rc != OPVD_OK) return rc;
        }
        if (max_avail + n - h->row_base > h->stride) return OPVD_ERR_CAPACITY;
    }
    bool uniform = true;
    for (int s = first; s < first + count; ++s) uniform = uniform && h->h_avail[s] == h->h_avail[first];
    if (uniform) {
        CK(cudaMemcpy2DAsync(h->d_iq_owned + (size_t)first * h->stride + (h->h_avail[first] - h->row_base),
                             (size_t)h->stride * 4, iq, (size_t)host_stride * 4, (size_t)n * 4, (size_t)count,
                             cudaMemcpyHostToDevice, h->st_copy));
    } else {
        for (int s = first; s < first + count; ++s)
            CK(cudaMemcpyAsync(h->d_iq_owned + (size_t)s * h->stride + (h->h_avail[s] - h->row_base),
                               iq + (size_t)(s - first) * host_stride * 2, (size_t)n * 4, cudaMemcpyHostToDevice, h->st_copy));
    }
    // the samples land in rows beyond what any enqueued kernel reads (they only append), so nothing on h->st has
    // to be waited for; the next opvd_run waits for this event before its kernels
    CK(cudaEventRecord(h->ev_copy, h->st_copy));
    h->copy_pending = true;
    for (int s = first; s < first + count; ++s) h->h_avail[s] += n;
    h->avail_dirty = true;
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
    CK(cudaSetDevice(h->dev));
    for (int s = 0; s < h->S; ++s) {
        const int64_t n = n_samples ? n_samples[s] : n_uniform;
        if (n < 0 || n > stride_samples) return OPVD_ERR_ARG;
        if (n < h->h_avail[s]) return OPVD_ERR_ARG;  // a stream never shrinks
        h->h_avail[s] = n;
    }
    h->d_iq = static_cast<const uint32_t*>(d_iq);
    h->stride = stride_samples;
    h->row_base = 0;
    h->attached = true;
    h->avail_dirty = true;
    return ensure_output_buffers(h, stride_samples);
}

int opvd_run(opvd_handle* h, int final_flag) {
    if (!h) return OPVD_ERR_ARG;
    if (!h->d_iq) return OPVD_ERR_STATE;
    CK(cudaSetDevice(h->dev));
    int rc = ensure_output_buffers(h, h->stride);
    if (rc != OPVD_OK) return rc;
    // capacity of the soft buffer for what this run can produce
    int64_t max_avail = 0;
    for (int s = 0; s < h->S; ++s) max_avail = std::max(max_avail, h->h_avail[s]);
    if (max_avail / kSps + 8 - h->soft_base > h->soft_stride) {
        if (h->cfg.mode == OPVD_MODE_STREAM && !h->attached) {
            rc = compact(h);
            if (rc != OPVD_OK) return rc;
        }
        if (max_avail / kSps + 8 - h->soft_base > h->soft_stride) return OPVD_ERR_CAPACITY;
    }
    if (h->avail_dirty) {
        CK(cudaMemcpyAsync(h->d_avail, h->h_avail.data(), sizeof(int64_t) * h->S, cudaMemcpyHostToDevice, h->st));
        CK(cudaStreamSynchronize(h->st));  // h_avail is pageable; keep it stable until the copy has landed
        h->avail_dirty = false;
    }
    if (h->copy_pending) {  // pushed samples must have landed before the kernels read them (device-side wait)
        CK(cudaStreamWaitEvent(h->st, h->ev_copy, 0));
        h->copy_pending = false;
    }
    StreamBuffers sb{h->d_iq, h->stride, h->d_avail, h->row_base};
    SoftBuffers so{h->d_soft, h->soft_stride, h->soft_base};
    CK(cudaMemsetAsync(h->d_ntasks, 0, sizeof(int32_t), h->st));
    CK(cudaEventRecord(h->ev[0], h->st));
    launch_estimate(sb, h->d_dstate, h->d_est, h->S, h->cfg.mode, final_flag ? 1 : 0, h->st);
    CK(cudaEventRecord(h->ev[1], h->st));
    if (h->cfg.coherent && h->cfg.mode == OPVD_MODE_BATCH) {
        CK(launch_demod_coherent(sb, so, h->d_dstate, h->S, final_flag ? 1 : 0, h->cfg.afc_alpha,
                                 h->cfg.pll_bw_hz > 0.0 ? h->cfg.pll_bw_hz : 50.0, h->d_counters, h->st));
    } else {
        CK(launch_demod(sb, so, h->d_dstate, h->S, h->cfg.mode, final_flag ? 1 : 0, h->cfg.afc_alpha,
                        h->cfg.lanes_per_stream, h->d_counters, h->st));
    }
    CK(cudaEventRecord(h->ev[2], h->st));
    launch_track(so, h->d_dstate, h->d_tstate, h->S, h->d_frec, h->max_frames, h->d_events, h->d_nevents,
                 h->max_events, h->d_tasks, h->d_ntasks, h->max_tasks, h->d_counters, h->st);
    CK(cudaEventRecord(h->ev[3], h->st));
    launch_decode(so, h->d_tasks, h->d_ntasks, h->max_tasks, h->d_frames, h->d_metrics, h->max_frames,
                  h->d_counters, h->st);
    CK(cudaEventRecord(h->ev[4], h->st));
    CK(cudaGetLastError());
    h->have_times = true;
    h->mirror_stale = true;
    h->ev_mirror_stale = true;
    if (final_flag) h->final_seen = true;
    return OPVD_OK;
}

int opvd_sync(opvd_handle* h) {
    if (!h) return OPVD_ERR_ARG;
    CK(cudaSetDevice(h->dev));
    CK(cudaStreamSynchronize(h->st_copy));
    CK(cudaStreamSynchronize(h->st));
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
    CK(cudaEventSynchronize(h->ev[4]));
    for (int i = 0; i < 4; ++i) CK(cudaEventElapsedTime(&ms5[i], h->ev[i], h->ev[i + 1]));
    CK(cudaEventElapsedTime(&ms5[4], h->ev[0], h->ev[4]));
    return OPVD_OK;
}

int opvd_poll_frames(opvd_handle* h, int32_t max_frames, uint8_t* frames134, opvd_frame_info* info) {
    if (!h || max_frames < 0 || (max_frames > 0 && !frames134)) return OPVD_ERR_ARG;
    CK(cudaSetDevice(h->dev));
    int rc = refresh_frame_mirror(h);
    if (rc != OPVD_OK) return rc;
    int n = 0;
    if (!h->d_frames) return 0;
    for (int s = 0; s < h->S && n < max_frames; ++s) {
        const int total = h->h_tstate[s].total_frames;
        if (total - h->polled_frames[s] > h->max_frames) return OPVD_ERR_CAPACITY;  // ring overrun
        while (h->polled_frames[s] < total && n < max_frames) {
            const int k = h->polled_frames[s]++;
            const size_t o = (size_t)s * h->max_frames + (k % h->max_frames);
            if (h->h_metrics[o] < 0) continue;  // dropped frame (:1052), never written by the reference
            memcpy(frames134 + (size_t)n * kFrameBytes, &h->h_frames[o * kFrameBytes], kFrameBytes);
            if (info) {
                opvd_frame_info& fi = info[n];
                fi.stream = s; fi.frame_idx = k; fi.metric = h->h_metrics[o]; fi.reserved = 0;
                fi.payload_start = h->h_frec[o].payload_start; fi.ready_idx = h->h_frec[o].ready_idx;
                fi.sync_quality = h->h_frec[o].quality;
            }
            ++n;
        }
        if (h->polled_frames[s] < total) break;  // output buffer full
    }
    return n;
}

int opvd_poll_events(opvd_handle* h, int32_t stream, int32_t max_events, opvd_event* out) {
    if (!h || stream < 0 || stream >= h->S || max_events < 0 || (max_events > 0 && !out)) return OPVD_ERR_ARG;
    if (!h->d_events) return 0;
    CK(cudaSetDevice(h->dev));
    if (h->ev_mirror_stale) {
        CK(cudaStreamSynchronize(h->st));
        h->h_nevents.resize(h->S);
        h->h_events.resize((size_t)h->S * h->max_events);
        CK(cudaMemcpy(h->h_nevents.data(), h->d_nevents, sizeof(int32_t) * h->S, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(h->h_events.data(), h->d_events, sizeof(TrackEvent) * h->h_events.size(), cudaMemcpyDeviceToHost));
        h->ev_mirror_stale = false;
    }
    const int total = h->h_nevents[stream];
    if (total - h->polled_events[stream] > h->max_events) return OPVD_ERR_CAPACITY;
    int n = 0;
    while (h->polled_events[stream] < total && n < max_events) {
        const int k = h->polled_events[stream]++;
        const TrackEvent& e = h->h_events[(size_t)stream * h->max_events + (k % h->max_events)];
        out[n].type = e.type; out[n].count = e.count; out[n].sym_idx = e.sym_idx; out[n].corr = e.corr; out[n].raw = e.raw;
        ++n;
    }
    return n;
}

int opvd_get_soft(opvd_handle* h, int32_t stream, int64_t first_sym, int64_t n, double* out) {
    if (!h || stream < 0 || stream >= h->S || first_sym < 0 || n < 0 || (n > 0 && !out)) return OPVD_ERR_ARG;
    if (!h->d_soft) return 0;
    CK(cudaSetDevice(h->dev));
    CK(cudaStreamSynchronize(h->st));
    DemodState ds;
    CK(cudaMemcpy(&ds, h->d_dstate + stream, sizeof(ds), cudaMemcpyDeviceToHost));
    if (first_sym < h->soft_base) return OPVD_ERR_ARG;
    const int64_t m = std::max<int64_t>(0, std::min<int64_t>(n, ds.n_sym - first_sym));
    if (m > 0)
        CK(cudaMemcpy(out, h->d_soft + (size_t)stream * h->soft_stride + (first_sym - h->soft_base), m * sizeof(double),
                      cudaMemcpyDeviceToHost));
    return (int)std::min<int64_t>(m, INT32_MAX);
}

int opvd_get_stream_info(opvd_handle* h, int32_t stream, opvd_stream_info* out) {
    if (!h || !out || stream < 0 || stream >= h->S) return OPVD_ERR_ARG;
    CK(cudaSetDevice(h->dev));
    CK(cudaStreamSynchronize(h->st));
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
    h->h_tstate.resize(h->S);
    CK(cudaMemcpy(h->h_tstate.data(), h->d_tstate, sizeof(TrackState) * h->S, cudaMemcpyDeviceToHost));
    std::vector<int32_t> nf(h->S);
    for (int s = 0; s < h->S; ++s) nf[s] = std::min(h->h_tstate[s].total_frames, h->max_frames);
    int32_t* d_nf = nullptr;
    CK(dalloc(&d_nf, h->S));
    CK(cudaMemcpy(d_nf, nf.data(), sizeof(int32_t) * h->S, cudaMemcpyHostToDevice));
    launch_bert_check_impl(h->d_frames, h->d_metrics, h->d_frec, d_nf, h->S, h->max_frames, to_params(p),
                           h->d_counters, h->st);
    CK(cudaStreamSynchronize(h->st));
    CK(cudaGetLastError());
    cudaFree(d_nf);
    return OPVD_OK;
}

}  // extern "C"
