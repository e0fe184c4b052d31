// tests/hostsim/hostsim.cpp — TEST INFRASTRUCTURE ONLY.
// Compiles the product's host/device arithmetic headers (demod_core.cuh, est_core.cuh,
// track_core.cuh) for the CPU and drives them one stream at a time, so that the restructured
// FP64 algorithm can be checked against the oracle in the authoring container (which has no GPU).
// It is never linked into libopvd / the product path.
#include <cstdint>
#include <cstring>
#include <vector>
#include <cmath>
#include "../../opv_cxx_demod_b200/csrc/demod_core.cuh"
#include "../../opv_cxx_demod_b200/csrc/est_core.cuh"
#include "../../opv_cxx_demod_b200/csrc/track_core.cuh"

using namespace opvd;

extern "C" {

double hostsim_estimate(const int16_t* iq, size_t n) {
    size_t test = n < (size_t)kEstSamples ? n : (size_t)kEstSamples;
    double Rr[kEstLags] = {0}, Ri[kEstLags] = {0};
    for (size_t blk = 0; blk < test / kSps; ++blk) {
        const int16_t* b = iq + 2 * blk * kSps;
        for (int l = 0; l < kEstLags; ++l)
            for (int i = 0; i + l < kSps; ++i) {
                double a = b[2 * i], bq = b[2 * i + 1], a2 = b[2 * (i + l)], b2 = b[2 * (i + l) + 1];
                Rr[l] += a2 * a + b2 * bq;
                Ri[l] += b2 * a - a2 * bq;
            }
    }
    return est_scan([&](double o) { return est_energy(Rr, Ri, o); });
}

// whole stream: returns number of soft symbols
size_t hostsim_demod(const int16_t* iq, size_t n, int mode, double afc_alpha, int have_init, double init_offset,
                     double* soft_out, size_t cap, double* est_out, double* final_freq, double* final_tfreq) {
    // packed samples with 64 words of readable padding in front (local indices < 0 of the first call)
    std::vector<uint32_t> w(n + 64 + 64, 0xDEADBEEFu);
    for (size_t i = 0; i < n; ++i)
        w[64 + i] = (uint32_t)(uint16_t)iq[2 * i] | ((uint32_t)(uint16_t)iq[2 * i + 1] << 16);
    const uint32_t* base = w.data() + 64;

    DemodState st;
    demod_state_init(st);
    double est = 0.0;
    if (mode == kModeBatch) {
        est = hostsim_estimate(iq, n);
        st.freq_offset = est;
    } else if (have_init) {
        st.freq_offset = init_offset;
    } else if (n >= (size_t)kChunkSamples) {
        est = hostsim_estimate(iq, kChunkSamples);
        st.freq_offset = est;
    }
    st.flags |= kFlagEstDone;
    DemodRegs r;
    regs_from_state(r, st);
    size_t ns = 0;
    while (demod_schedule(st, r.pos, mode, (int64_t)n, true)) {
        const int64_t b = (int64_t)r.pos;
        const double f = r.pos - (double)b;
        const uint32_t* win = base + st.origin + b - kWinLead;
        double soft = demod_symbol(r, win, f, st.sym_in_call == 0, afc_alpha);
        st.sym_in_call++;
        if (ns < cap) soft_out[ns] = soft;
        ++ns;
        st.n_sym++;
    }
    regs_to_state(r, st);
    if (est_out) *est_out = est;
    if (final_freq) *final_freq = st.freq_offset;
    if (final_tfreq) *final_tfreq = st.timing_freq;
    return ns;
}

// scalar event-driven tracker over a soft row; returns number of frame records
size_t hostsim_track(const double* soft, size_t n_sym, FrameRec* frames, size_t cap_frames,
                     TrackEvent* events, size_t cap_events, size_t* n_events_out, int* final_state) {
    TrackState t;
    track_state_init(t);
    size_t nf = 0, ne = 0;
    auto ev = [&](int type, int count, int64_t idx, double corr, double raw) {
        if (ne < cap_events) events[ne] = {type, count, idx, corr, raw};
        ++ne;
    };
    const int64_t N = (int64_t)n_sym;
    for (;;) {
        if (t.state == kHunting) {
            int64_t n = t.cursor < kSyncBits - 1 ? kSyncBits - 1 : t.cursor;
            bool hit = false;
            for (; n < N; ++n) {
                double raw, norm = sync_correlate(soft + n - (kSyncBits - 1), raw);
                if (hunt_hit(norm, raw)) {
                    t.state = kVerifying; t.quality = norm; t.anchor = n; t.collecting = 1; t.payload_start = n + 1;
                    ev(kEvHuntToVerify, 0, n, norm, raw);
                    hit = true;
                    break;
                }
            }
            if (!hit) { t.cursor = N; break; }
        } else if (t.state == kVerifying) {
            const int64_t ready = t.anchor + kEncodedBits;
            if (ready >= N) break;
            if (nf < cap_frames) frames[nf] = {t.payload_start, ready, t.quality};
            ++nf;
            t.total_frames++; t.collecting = 0; t.state = kLocked; t.misses = 0;
            ev(kEvVerifyToLocked, t.total_frames, ready, 0, 0);
        } else {
            if (t.collecting) {
                const int64_t ready = t.payload_start + kEncodedBits - 1;
                if (ready >= N) break;
                if (nf < cap_frames) frames[nf] = {t.payload_start, ready, t.quality};
                ++nf;
                t.total_frames++; t.collecting = 0;
            }
            const int64_t nb = t.anchor + kFrameSymbols;
            if (nb >= N) break;
            double raw, corr = sync_correlate(soft + nb - (kSyncBits - 1), raw);
            if (corr >= 0.70) {
                t.misses = 0; t.quality = corr; t.collecting = 1; t.payload_start = nb + 1; t.anchor = nb;
                ev(kEvSyncOk, 0, nb, corr, raw);
            } else {
                t.misses++;
                ev(kEvSyncMiss, t.misses, nb, corr, raw);
                if (t.misses >= kSyncMissLimit) {
                    t.state = kHunting; t.collecting = 0; t.cursor = nb + 1;
                    ev(kEvLostLock, 0, nb, 0, 0);
                } else {
                    t.quality = corr; t.collecting = 1; t.payload_start = nb + 1; t.anchor = nb;
                }
            }
        }
    }
    if (n_events_out) *n_events_out = ne;
    if (final_state) *final_state = t.state;
    return nf;
}

}  // extern "C"
