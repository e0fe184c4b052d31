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
#include "../../opv_cxx_demod_b200/csrc/demod_warp_core.cuh"
#include "../../opv_cxx_demod_b200/csrc/demod_bank_core.cuh"
#include "../../opv_cxx_demod_b200/csrc/demod_coherent_core.cuh"
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

// whole stream through the WARP-PER-STREAM lane decomposition (demod_warp_core.cuh): the 32 lanes are
// simulated in lock step, warp shuffles become array indexing.  Same contract as hostsim_demod.
size_t hostsim_demod_warp(const int16_t* iq, size_t n, int mode, double afc_alpha, int have_init, double init_offset,
                          double* soft_out, size_t cap, double* est_out, double* final_freq, double* final_tfreq) {
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

    WarpLane wl[32];
    for (int l = 0; l < 32; ++l) {
        warp_lane_init(wl[l], l, g_fm);
        warp_lane_lo(wl[l], st.freq_offset);
        wl[l].prev = wl[l].tone ? st.p2 : st.p1;
    }
    cplx RP[32];
    bool prev_zero[32];
    for (int l = 0; l < 32; ++l) {
        RP[l] = cmul(wl[l].R, wl[l].prev);
        prev_zero[l] = wl[l].prev.r == 0.0 && wl[l].prev.i == 0.0;
    }
    double freq_offset = st.freq_offset, pos = st.pos, timing_freq = st.timing_freq, ph1 = st.ph1, ph2 = st.ph2;
    auto down = [](const cplx* v, int l, int d) { return (l + d < 32) ? v[l + d] : v[l]; };
    size_t ns = 0;
    while (demod_schedule(st, pos, mode, (int64_t)n, true)) {
        const int64_t b = (int64_t)pos;
        const double f = pos - (double)b;
        const uint32_t* win = base + st.origin + b - kWinLead;
        const bool first = st.sym_in_call == 0;
        cplx W[32], F[32], A[32], B[32], Cc[32], X[32];
        double nrm[32], pdo[32];
        for (int l = 0; l < 32; ++l) {
            const int pc = wl[l].p > 12 ? 12 : wl[l].p;
            uint32_t s5[5];
            for (int r = 0; r < 5; ++r) s5[r] = (5 * pc + r <= 60) ? win[5 * pc + r] : 0x12345678u;  // beyond the window: junk, unused
            LanePartial lp = warp_lane_partial(wl[l], s5);
            W[l] = lp.W; F[l] = lp.F;
        }
        for (int l = 0; l < 32; ++l) { cplx o = down(W, l, 1); A[l] = {W[l].r + o.r, W[l].i + o.i}; }
        for (int l = 0; l < 32; ++l) { cplx o = down(A, l, 2); B[l] = {A[l].r + o.r, A[l].i + o.i}; }
        for (int l = 0; l < 32; ++l) { cplx o = down(B, l, 4); Cc[l] = {B[l].r + o.r, B[l].i + o.i}; }
        for (int l = 0; l < 32; ++l) X[l] = warp_lane_gate(wl[l], f, Cc[l], down(F, l, 8), F[l]);
        if (first) {
            auto acc = [&](int k) { return win[k]; };
            for (int l = 0; l < 32; l += 16) {
                const cplx fix = first_symbol_fix_w(acc, f, wl[l].z);
                X[l].r -= fix.r; X[l].i -= fix.i;
            }
        }
        for (int l = 0; l < 32; ++l) nrm[l] = cnorm(X[l]);
        bool tone1;
        const double soft = warp_uniform_timing(nrm[2], nrm[18], nrm[0], nrm[4], nrm[16], nrm[20], timing_freq, pos, tone1, g_fm);
        for (int l = 0; l < 32; ++l) {
            const bool xz = nrm[l] == 0.0;
            pdo[l] = first ? 0.0 : warp_lane_afc_phase(wl[l], X[l], RP[l], xz || prev_zero[l], wl[l].tone ? ph2 : ph1, g_fm);
            prev_zero[l] = xz;
        }
        if (!first) warp_afc_loop(freq_offset, pdo[tone1 ? 2 : 18], afc_alpha, g_fm);
        for (int l = 0; l < 32; ++l) wl[l].prev = cmul(X[l], cconj(wl[(l & 16) | 10].R));
        ph1 = warp_wrap_phase(fma(40.0, wl[0].inc, ph1), g_fm);
        ph2 = warp_wrap_phase(fma(40.0, wl[16].inc, ph2), g_fm);
        if (!first) for (int l = 0; l < 32; ++l) warp_lane_lo_fast(wl[l], freq_offset, g_fm);
        for (int l = 0; l < 32; ++l) RP[l] = cmul(wl[l].R, wl[l].prev);
        st.sym_in_call++;
        if (ns < cap) soft_out[ns] = soft;
        ++ns;
        st.n_sym++;
    }
    if (est_out) *est_out = est;
    if (final_freq) *final_freq = freq_offset;
    if (final_tfreq) *final_tfreq = timing_freq;
    return ns;
}

// whole stream through the CHANNEL-BANK decomposition (demod_bank_core.cuh, kernels_demod_bank.cu): on-time sums of
// both tones, early/late sums of the dominant tone only, LO powers from one zeta chain.  Same contract as hostsim_demod.
static size_t demod_bank_impl(bool elb, const int16_t* iq, size_t n, int mode, double afc_alpha, int have_init, double init_offset,
                              double* soft_out, size_t cap, double* est_out, double* final_freq, double* final_tfreq) {
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
    BankAfc afc = {st.freq_offset, st.ph1, st.ph2, st.p1, st.p2};
    BankLo lo;
    BankPow pw;
    {
        double d;
        const cplx zeta = bank_zeta_general(afc.freq_offset, d);
        bank_lo_from_zeta(zeta, d, lo, g_fm);
        bank_pow_from_zeta(zeta, pw, g_bk);
    }
    double pos = st.pos, timing_freq = st.timing_freq;
    size_t ns = 0;
    while (demod_schedule(st, pos, mode, (int64_t)n, true)) {
        const int64_t b = (int64_t)pos;
        const double f = pos - (double)b;
        const uint32_t* win = base + st.origin + b - kWinLead;
        const bool first = st.sym_in_call == 0;
        auto slot = [&](int k, double& I, double& Q) { unpack_iq(win[k], I, Q); };
        cplx A[4], B[4], s10, s20, s40;
        bank_on_blocks(slot, lo.z1, lo.z2, A, B, s10, s20, s40);
        BankOnTime on;
        bank_on_time(slot, f, lo, pw, A, B, s10, s20, s40, on);
        const bool tone1 = on.eO1 > on.eO2;
        const double soft = on.eO2 - on.eO1;
        cplx fixE = {0.0, 0.0};
        if (first) fixE = first_symbol_fix(win, f, tone1 ? lo.z1 : lo.z2);
        double eE, eL;
        if (elb) {  // the AFC warp evaluates H0, H5 of both tones, the window warp combines the dominant tone's
            BankElBlocks e;
            bank_el_blocks(slot, lo.z1, lo.z2, e);
            bank_early_late_from_blocks(f, tone1, lo, pw, on, tone1 ? e.H0a : e.H0b, tone1 ? e.H5a : e.H5b, e.s0, e.s60, fixE,
                                        eE, eL);
        } else {
            bank_early_late(slot, f, tone1, lo, pw, on, fixE, eE, eL);
        }
        bank_timing(eE, eL, timing_freq, pos, g_fm);
        // AFC role
        bank_afc(afc, on.O1, on.O2, tone1, pw.zeta40, lo.inc1, lo.inc2, first, afc_alpha, g_fm);
        if (!first) {
            double d;
            const cplx zeta = bank_zeta_fast(afc.freq_offset, d, g_fm);
            bank_lo_from_zeta(zeta, d, lo, g_fm);
            bank_pow_from_zeta(zeta, pw, g_bk);
        }
        st.sym_in_call++;
        if (ns < cap) soft_out[ns] = soft;
        ++ns;
        st.n_sym++;
    }
    if (est_out) *est_out = est;
    if (final_freq) *final_freq = afc.freq_offset;
    if (final_tfreq) *final_tfreq = timing_freq;
    return ns;
}


size_t hostsim_demod_bank(const int16_t* iq, size_t n, int mode, double afc_alpha, int have_init, double init_offset,
                          double* soft_out, size_t cap, double* est_out, double* final_freq, double* final_tfreq) {
    return demod_bank_impl(false, iq, n, mode, afc_alpha, have_init, init_offset, soft_out, cap, est_out, final_freq, final_tfreq);
}
// the same with the early / late block sums taken by the AFC role (the kernel's choice for banks of up to one CTA per SM)
size_t hostsim_demod_bank_elb(const int16_t* iq, size_t n, int mode, double afc_alpha, int have_init, double init_offset,
                              double* soft_out, size_t cap, double* est_out, double* final_freq, double* final_tfreq) {
    return demod_bank_impl(true, iq, n, mode, afc_alpha, have_init, init_offset, soft_out, cap, est_out, final_freq, final_tfreq);
}

// coherent mode (-c, batch only) through demod_coherent_core.cuh.  Returns number of soft symbols.
size_t hostsim_demod_coherent(const int16_t* iq, size_t n, double afc_alpha, double pll_bw, double* soft_out, size_t cap,
                              double* est_out, double* final_freq) {
    const double est = hostsim_estimate(iq, n);
    CoherentState st;
    coherent_init(st, est, afc_alpha, pll_bw);
    size_t ns = 0;
    for (size_t sym = 0; sym < n / kSps; ++sym) {
        double I[kSps], Q[kSps];
        for (int i = 0; i < kSps; ++i) { I[i] = iq[2 * (sym * kSps + i)]; Q[i] = iq[2 * (sym * kSps + i) + 1]; }
        const double soft = coherent_symbol(st, I, Q, sym == 0);
        if (ns < cap) soft_out[ns] = soft;
        ++ns;
    }
    if (est_out) *est_out = est;
    if (final_freq) *final_freq = st.freq_offset;
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
