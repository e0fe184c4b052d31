/* oracle/opv_oracle.h — TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the opv-demod receive chain (and of the opv-mod
 * transmit chain, used only to make captures).  It exists to CHECK the CUDA
 * product; nothing under opv_cxx_demod_b200/ may include, link or call it.
 * Allowed callers: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline /
 * --impl reference legs.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_ref.py checks every function here
 * bit-for-bit against the reference itself (oracle/_ref/, built from
 * /root/reference/src by oracle/Makefile), and tests/golden/ holds hashes of
 * reference outputs produced in the authoring container.
 *
 * Citations are to /root/reference/src/opv-demod.cpp unless prefixed "mod:"
 * (= /root/reference/src/opv-mod.cpp).
 */
#ifndef OPV_ORACLE_H
#define OPV_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    ORA_SPS = 40,             /* :39  */
    ORA_SYNC_BITS = 24,       /* :47  */
    ORA_FRAME_BYTES = 134,    /* :49  */
    ORA_FRAME_BITS = 1072,    /* :50  */
    ORA_ENCODED_BITS = 2144,  /* :51  */
    ORA_FRAME_SYMBOLS = 2168, /* :52  */
    ORA_CHUNK_SAMPLES = 86720 /* :1012 */
};

/* ---- A2/A3: MSKDemodulatorAFC (:108-348) ---- */
typedef struct {
    double freq_offset, phase_f1, phase_f2;
    double prev1_re, prev1_im, prev2_re, prev2_im;
    double afc_alpha, mu, timing_freq, alpha_timing, beta_timing;
    size_t leftover;
} ora_demod_t;

void ora_demod_init(ora_demod_t* d);
double ora_estimate_offset(const int16_t* iq, size_t n);
/* one demodulate() call (:206-329); returns number of soft symbols (may exceed cap; only cap are stored) */
size_t ora_demodulate(ora_demod_t* d, const int16_t* iq, size_t n, double* soft_out, size_t cap);

/* ---- CoherentMSKDemodulator (:365-572), the batch-only `-c` alternative ---- */
typedef struct {
    double freq_offset, carrier_phase, phase_f1, phase_f2, loop_freq;
    double prev_re, prev_im;
    double afc_alpha, pll_alpha, pll_beta;
} ora_codemod_t;
void ora_codemod_init(ora_codemod_t* d);                       /* :367-377 */
void ora_codemod_set_pll_bandwidth(ora_codemod_t* d, double bw); /* :561-568 */
size_t ora_codemod_demodulate(ora_codemod_t* d, const int16_t* iq, size_t n, double* soft_out, size_t cap); /* :450-548 */

/* ---- A5: SyncTracker (:587-787) ---- */
enum { ORA_HUNTING = 0, ORA_VERIFYING = 1, ORA_LOCKED = 2 };
enum {
    ORA_EV_HUNT_TO_VERIFY = 1, /* :651 */
    ORA_EV_VERIFY_TO_LOCKED = 2, /* :677 */
    ORA_EV_SYNC_OK = 3,        /* :695 */
    ORA_EV_SYNC_MISS = 4,      /* :699 */
    ORA_EV_LOST_LOCK = 5       /* :705 */
};
typedef struct {
    int32_t type;
    int32_t count;   /* frame number (type 2) or miss number (type 4), else 0 */
    int64_t sym_idx;
    double corr;     /* normalised correlation (types 1,3,4) */
    double raw;      /* raw correlation (type 1) */
} ora_event_t;

typedef struct {
    int state;
    double corr_buf[ORA_SYNC_BITS];
    size_t corr_idx;
    size_t total_symbols;
    int collecting;
    double pending[ORA_ENCODED_BITS];
    size_t n_pending;
    size_t since_sync;
    double sync_quality;
    int misses;
    int total_frames;
} ora_tracker_t;

void ora_tracker_init(ora_tracker_t* t);
/* returns 1 when a frame payload is ready (copied to payload_out[2144]); appends at most 2 events */
int ora_tracker_process(ora_tracker_t* t, double soft, size_t sym_idx, double* payload_out,
                        double* quality_out, ora_event_t* ev, int* n_ev);

/* ---- A6-A8: FrameDecoder / ViterbiDecoder (:792-902) ---- */
size_t ora_deinterleave_addr(size_t idx);
int ora_viterbi_decode(const int32_t* soft_in2144, uint8_t* bits1072);
int ora_frame_decode(const double* soft2144, uint8_t* out134);
void ora_quantise(const double* soft2144, int32_t* q2144, double* scale_out); /* :856-866, no deinterleave */
void ora_lfsr_table(uint8_t* out134);
void ora_set_codemod_perturb(double rad); /* test hook, see opv_oracle.c */

/* ---- whole chain: main() drivers (:995-1216) ---- */
typedef struct {
    int streaming;        /* -s */
    double afc_alpha;     /* -a, default 0.001 */
    int have_init_offset; /* -o (honoured only with -s, :1004 vs :1164) */
    double init_offset;
    int coherent;         /* -c (batch only: the streaming branch returns first, :995-1125) */
    double pll_bw;        /* -p, default 50.0 (:946) */
} ora_cfg_t;

typedef struct {
    /* caller-provided capacity */
    uint8_t* frames; int32_t* metrics; int64_t* frame_ready_idx; size_t cap_frames;
    double* soft; size_t cap_soft;
    ora_event_t* events; size_t cap_events;
    int64_t* chunk_starts; size_t cap_chunks;
    /* results */
    size_t n_frames;      /* frames with metric >= 0 ("decoded", :1053) */
    size_t n_perfect;     /* metric == 0 */
    size_t n_dropped;     /* metric < 0 (scale < 1e-10, :859) */
    size_t n_soft, n_events, n_chunks;
    double est_offset, final_freq, final_tfreq;
    int final_state;
    size_t total_samples; /* as summed by the streaming driver (:1027), 0 in batch */
} ora_result_t;

int ora_run(const ora_cfg_t* cfg, const int16_t* iq, size_t n_samples, ora_result_t* res);

/* ---- TX chain restatement (capture generator; mod:59-361) ---- */
void ora_base40_encode(const char* callsign, uint8_t* out6);
void ora_bert_frame(const char* callsign, uint32_t token, uint32_t frame_num, uint8_t* out134);
void ora_encode_frame(const uint8_t* payload134, uint8_t* bits2144);
/* modulate n_frames 134-byte frames (sync + payload each) followed by 4000 zero samples.
 * out must hold (n_frames*2168*40 + 4000) I/Q pairs.  Returns number of samples. */
size_t ora_modulate_frames(const uint8_t* frames, size_t n_frames, int16_t* out_iq);

#ifdef __cplusplus
}
#endif
#endif
