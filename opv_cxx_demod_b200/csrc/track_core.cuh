// track_core.cuh — 24-bit soft sync correlator and the HUNTING / VERIFYING / LOCKED flywheel
// (reference: SyncTracker, /root/reference/src/opv-demod.cpp:587-787) expressed over absolute
// symbol indices so that a stream is advanced event by event instead of symbol by symbol:
//   HUNTING   test every symbol n >= 23 (first hit wins, :637-653)
//   VERIFYING frame ready at n_hit + 2144 (:657-680)
//   LOCKED    one correlation at n_prev + 2168 (:684), flywheel for up to 4 misses, the 5th returns
//             to HUNTING (:696-713); the payload that started at the boundary is ready 2144 later (:720).
// The correlation itself keeps the reference's sequential summation order, so with identical soft
// symbols every threshold decision is identical.
#pragma once
#include "opvd_common.cuh"

namespace opvd {

struct TrackState {
    int32_t state;         // SyncState
    int32_t misses;        // consecutive_misses_
    int32_t total_frames;  // total_frames_ (frame_ready count, including frames the decoder drops)
    int32_t collecting;    // collecting_payload_
    int64_t cursor;        // HUNTING: next symbol index to test
    int64_t anchor;        // VERIFYING: hit index; LOCKED: index the next boundary is counted from
    int64_t payload_start; // first symbol of the payload being collected
    double quality;        // sync_quality_
};

OPVD_HD void track_state_init(TrackState& t) {
    t.state = kHunting; t.misses = 0; t.total_frames = 0; t.collecting = 0;
    t.cursor = 0; t.anchor = 0; t.payload_start = 0; t.quality = 0.0;
}

struct TrackEvent {
    int32_t type;
    int32_t count;
    int64_t sym_idx;
    double corr;
    double raw;
};

struct FrameRec {
    int64_t payload_start;  // soft[payload_start .. payload_start+2143]
    int64_t ready_idx;      // symbol index at which the reference reports frame_ready
    double quality;
};

// soft_correlate (:743-757) for the window ending at symbol n: w points at soft[n-23]
OPVD_HD double sync_correlate(const double* w, double& raw) {
    double sum = 0.0, energy = 0.0;
#pragma unroll
    for (int i = 0; i < kSyncBits; ++i) {
        const double s = w[i];
        const bool one = (kSyncWord >> (kSyncBits - 1 - i)) & 1u;  // bit 1 -> pattern -1 (:597-600)
        sum += one ? -s : s;                                     // s * (+/-1.0) is exact
        energy += fabs(s);
    }
    raw = sum;
    if (energy < 100.0) return 0.0;
    return sum / energy;
}

OPVD_HD bool hunt_hit(double norm, double raw) { return raw >= 5000.0 && norm >= 0.85; }  // :642

}  // namespace opvd
