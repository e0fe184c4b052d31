/* include/opvd.h — C ABI of the B200-native opv-demod receive chain (libopvd.so).
 *
 * The reference (OpenResearchInstitute/opv-cxx-demod) has NO library interface for this path: its
 * only boundary is the opv-demod process contract (argv, int16 LE I/Q on stdin, 134-byte frames
 * on stdout; /root/reference/src/opv-demod.cpp:943-1217, callers src/opv-modem.cpp:391,714 and
 * scripts/opv-pluto-rx.sh:286-288).  This header is therefore the interface a maintainer would
 * bind if the reference's objects were exported; each entry point names the reference seam it
 * replaces.  The drop-in executable (opv_cxx_demod_b200/bin/opv-demod) is a thin main() over it.
 *
 * Conventions: plain pointers and sizes, caller-owned buffers, int return codes (0 = OK, < 0 =
 * error, see opvd_strerror), no exceptions cross the ABI, one host thread per handle, one handle
 * per GPU.  All compute runs in hand-written sm_100a CUDA kernels; there is no CPU fallback: every
 * entry point fails with OPVD_ERR_CUDA when no usable device is present.
 */
#ifndef OPVD_H
#define OPVD_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OPVD_FRAME_BYTES 134      /* src/opv-demod.cpp:49 */
#define OPVD_FRAME_SYMBOLS 2168   /* :52 */
#define OPVD_ENCODED_BITS 2144    /* :51 */
#define OPVD_SPS 40               /* :39 */
#define OPVD_CHUNK_SAMPLES 86720  /* :1012 */

enum { OPVD_MODE_BATCH = 0,   /* opv-demod without -s: load all, one demodulate() call (:1127-1216) */
       OPVD_MODE_STREAM = 1   /* opv-demod -s: 86,720-sample chunks with leftover carry (:995-1125) */ };

enum { OPVD_OK = 0, OPVD_ERR_ARG = -1, OPVD_ERR_CUDA = -2, OPVD_ERR_CAPACITY = -3, OPVD_ERR_STATE = -4,
       OPVD_ERR_ALIGN = -5 };

enum { OPVD_HUNTING = 0, OPVD_VERIFYING = 1, OPVD_LOCKED = 2 };  /* enum class SyncState, :73 */

/* sync events = the lines SyncTracker prints on stderr (:651,:677,:695,:699,:705) */
enum { OPVD_EV_HUNT_TO_VERIFY = 1, OPVD_EV_VERIFY_TO_LOCKED = 2, OPVD_EV_SYNC_OK = 3, OPVD_EV_SYNC_MISS = 4,
       OPVD_EV_LOST_LOCK = 5 };

typedef struct opvd_handle opvd_handle;

typedef struct opvd_config {
    int32_t n_streams;         /* independent 2.168 MSPS streams handled by this handle (this GPU) */
    int32_t mode;              /* OPVD_MODE_* */
    double afc_alpha;          /* -a <alpha>, default 0.001 (:945,:955) */
    int32_t have_init_offset;  /* -o given; honoured in stream mode only, skips the estimate (:1004,:1031) */
    int32_t device;            /* CUDA device ordinal, -1 = current device */
    double init_offset_hz;     /* -o <hz> */
    int64_t max_samples;       /* capacity (samples per stream) of the library-owned I/Q buffer used by
                                  opvd_push_iq*; 0 when captures are attached with opvd_attach_device_iq.  Batch mode: the
                                  whole capture must fit.  Stream mode: the buffer is a ring (nothing is ever moved); it must
                                  hold one push plus one chunk (86,720 samples) of carry, and two pushes for a push to
                                  overlap the kernels of the previous run */
    int64_t max_symbols;       /* soft-symbol buffer capacity per stream (a ring in stream mode); 0 = derived from
                                  max_samples */
    int32_t max_frames;        /* frames per stream that may wait between two polls (the device frame log holds
                                  n_streams * max_frames entries); 0 = derived */
    int32_t lanes_per_stream;  /* demodulator kernel variant: 0 = chosen from n_streams; 32 = one warp per stream (small
                                  banks, lowest per-symbol latency); 96 = channel-bank kernel, 32 streams per CTA with
                                  three role warps (large banks) */
    int32_t coherent;          /* -c: CoherentMSKDemodulator (:365-572) instead of MSKDemodulatorAFC; honoured in batch
                                  mode only, like the reference (the streaming branch returns first, :995-1125) */
    int32_t reserved0;
    double pll_bw_hz;          /* -p <hz>, Costas loop bandwidth (coherent only), used verbatim like the reference (-p 0
                                  freezes the loop, :1149); NaN selects the reference's default 50.0 (:946) */
} opvd_config;

typedef struct opvd_event {
    int32_t type;     /* OPVD_EV_* */
    int32_t count;    /* frame number (VERIFY_TO_LOCKED) or miss number (SYNC_MISS) */
    int64_t sym_idx;  /* global symbol index printed by the reference in [brackets] */
    double corr;      /* normalised correlation */
    double raw;       /* raw correlation (HUNT_TO_VERIFY) */
} opvd_event;

typedef struct opvd_frame_info {
    int32_t stream;
    int32_t frame_idx;      /* 0-based index among this stream's frame_ready events */
    int32_t metric;         /* Viterbi path metric (0 = "perfect", :1054) */
    int32_t reserved;
    int64_t payload_start;  /* global symbol index of the first payload symbol (sync ended one before) */
    int64_t ready_idx;      /* global symbol index at which the reference reports the frame */
    double sync_quality;    /* res.sync_quality (:661,:722) */
} opvd_frame_info;

typedef struct opvd_stream_info {
    double est_offset_hz;   /* result of estimate_offset (0 when it never ran) */
    double freq_offset_hz;  /* demod.get_freq_offset() (:331) */
    double timing_freq;     /* demod.get_timing_freq() (:332) */
    int64_t n_symbols;      /* total soft symbols */
    int64_t n_samples_used; /* call origin (samples consumed by completed calls) */
    int32_t sync_state;     /* tracker.get_state() (:738) */
    int32_t frames_ready;   /* tracker.get_total_frames() (:739) */
    int32_t done;           /* EOF processing finished */
    int32_t reserved;
} opvd_stream_info;

enum { OPVD_CTR_SAMPLES = 0, OPVD_CTR_SYMBOLS, OPVD_CTR_FRAMES_READY, OPVD_CTR_FRAMES_DECODED, OPVD_CTR_FRAMES_PERFECT,
       OPVD_CTR_FRAMES_DROPPED, OPVD_CTR_SYNC_ACQ, OPVD_CTR_SYNC_OK, OPVD_CTR_SYNC_MISS, OPVD_CTR_LOST_LOCK,
       OPVD_CTR_BIT_ERRORS, OPVD_CTR_FRAMES_COMPARED, OPVD_CTR_ACS, OPVD_NUM_COUNTERS = 16 };

/* ---- lifetime.  Replaces the three objects main() constructs per process
 *      (MSKDemodulatorAFC demod; SyncTracker tracker; FrameDecoder fdec; :999-1001 / :1164,:1182-1183),
 *      one set per stream. */
int opvd_create(const opvd_config* cfg, opvd_handle** out);
int opvd_destroy(opvd_handle* h);
/* back to the state right after opvd_create (fresh demod/tracker/decoder objects, counters zero); buffers and an
 * attached capture are kept.  Lets a benchmark repeat the same pass without reallocating. */
int opvd_reset(opvd_handle* h);
const char* opvd_strerror(int code);
const char* opvd_last_cuda_error(const opvd_handle* h);
int opvd_version(void);

/* ---- input.  Replaces the stdin read loops (:1022-1023 streaming, :1132-1135 batch).
 * Samples are interleaved int16 I,Q (4 bytes each), host endian, exactly the reference's stdin bytes. */
/* append n_samples to ONE stream from a host buffer (pinned memory makes the copy asynchronous) */
int opvd_push_iq(opvd_handle* h, int32_t stream, const int16_t* iq, int64_t n_samples);
/* append n_samples to EVERY stream from a host array laid out [n_streams][host_stride_samples] */
int opvd_push_iq_all(opvd_handle* h, const int16_t* iq, int64_t n_samples, int64_t host_stride_samples);
/* use captures already resident in device memory: [n_streams][stride_samples] packed I/Q words.
 * d_iq must be 16-byte aligned and stride_samples a multiple of 4 (TMA bulk-copy granularity).
 * n_samples: per-stream valid lengths (host array) or NULL for n_uniform everywhere.  May be called again with
 * larger lengths for the same buffer (more of the capture becomes visible: time tiles over a resident bank); after
 * opvd_reset any lengths are accepted again. */
int opvd_attach_device_iq(opvd_handle* h, const void* d_iq, int64_t stride_samples, const int64_t* n_samples,
                          int64_t n_uniform);

/* ---- processing.  Replaces estimate_offset + demodulate + tracker.process + fdec.decode for all
 * streams (:1030-1065, :1166-1205).  final != 0 is EOF: batch mode runs its single call, stream mode
 * flushes the remainder (:1088-1113).  Work is only enqueued (estimate + demodulate on one CUDA stream,
 * tracker + Viterbi on a second one, so they overlap the demodulator of the next run); the call never
 * waits for the device.  OPVD_ERR_CAPACITY: the run could produce more soft symbols than max_symbols holds. */
int opvd_run(opvd_handle* h, int final_flag);
int opvd_sync(opvd_handle* h);

/* ---- output.  Replaces cout.write(frame,134) (:1059-1062, :1200-1203).
 * Waits for the runs enqueued so far, then returns the number of frames written (<= max_frames), frames with
 * metric >= 0 only; each frame is returned once, in stream order within a poll and in frame order within a
 * stream.  Only frames decoded since the last poll cross PCIe.  info may be NULL. */
int opvd_poll_frames(opvd_handle* h, int32_t max_frames, uint8_t* frames134, opvd_frame_info* info);
/* the same without waiting: the frames of the runs that have already finished (a live ingest loop polls with this
 * between pushes so that the host never idles the copy engine; the reference writes its frames as they appear too) */
int opvd_poll_frames_ready(opvd_handle* h, int32_t max_frames, uint8_t* frames134, opvd_frame_info* info);
/* frames that were overwritten in the device log before anybody polled them (more than n_streams * max_frames
 * frames between two polls); polling continues with the oldest surviving frame */
int opvd_frames_lost(opvd_handle* h, uint64_t* out);
/* events of one stream not yet returned (the tracker's stderr lines) */
int opvd_poll_events(opvd_handle* h, int32_t stream, int32_t max_events, opvd_event* out);
/* soft symbols [first_sym, first_sym+n) of one stream (parity/debug; the reference never exposes them) */
int opvd_get_soft(opvd_handle* h, int32_t stream, int64_t first_sym, int64_t n, double* out);
int opvd_get_stream_info(opvd_handle* h, int32_t stream, opvd_stream_info* out);
/* this rank's counters (OPVD_CTR_*); reduce across ranks with NCCL/torch.distributed */
int opvd_get_counters(opvd_handle* h, uint64_t* out, int32_t n);
/* device pointer of the same counters (uint64[OPVD_NUM_COUNTERS]) for an in-place ncclAllReduce */
int opvd_counters_device_ptr(opvd_handle* h, void** out);
/* GPU time since the last opvd_create / opvd_reset, measured with CUDA events on the handle's streams:
 * ms[0..3] = summed kernel times of the estimate, demodulator, tracker and decoder launches of every run,
 * ms[4] = elapsed time from the start of the first run to the end of the last one (with several runs in flight
 * the tracker + decoder of run t overlap the demodulator of run t+1, so ms[4] < ms[0]+..+ms[3]).  Waits for the
 * enqueued runs. */
int opvd_last_run_ms(opvd_handle* h, float* ms5);
/* the demodulator kernel variant in use (opvd_config.lanes_per_stream with 0 resolved) */
int opvd_demod_lanes(opvd_handle* h);

/* ---- stage-level seams (parity tests) */
/* FrameDecoder::decode (:854-898): n payloads of 2144 doubles -> n frames of 134 bytes + metrics (-1 = dropped) */
int opvd_stage_decode(int32_t device, const double* payloads, int32_t n, uint8_t* frames134, int32_t* metrics);
/* same, device pointers, no host copies, returns elapsed ms in *ms (may be NULL) */
int opvd_stage_decode_dev(int32_t device, const double* d_payloads, int32_t n, uint8_t* d_frames134,
                          int32_t* d_metrics, float* ms);

/* ---- synthetic channel bank (measurement aid, not part of the reference's receive path) */
typedef struct opvd_synth {
    int32_t n_streams, n_frames;
    int64_t stride_samples, n_samples;
    uint64_t seed;
    float scale, ebn0_lo_db, ebn0_hi_db, cfo_max_hz;
    int32_t frac_delay, max_lead, first_stream, reserved;
} opvd_synth;
int opvd_synth_bank(int32_t device, const opvd_synth* p, void* d_iq);
/* compare every decoded frame of the handle with the bank's known BERT payloads -> BIT_ERRORS / FRAMES_COMPARED */
int opvd_bert_check(opvd_handle* h, const opvd_synth* p);

#ifdef __cplusplus
}
#endif
#endif
