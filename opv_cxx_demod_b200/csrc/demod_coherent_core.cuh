// demod_coherent_core.cuh — per-symbol arithmetic of the coherent (Costas-loop) MSK demodulator, the
// reference's batch-only `-c` alternative (CoherentMSKDemodulator::demodulate,
// /root/reference/src/opv-demod.cpp:450-548; selected at :1144-1161).
// Host/device code: the CUDA kernel (kernels_demod_coherent.cu) and the CPU host-sim test compile
// exactly these functions.
//
// What the reference computes per symbol (fixed 40-sample grid, no timing recovery):
//     corr_t = sum_{i<40} s[i] * exp(-j(cp + i*lf)) * exp(-j(ph_t + i*inc_t))
// with cp = carrier_phase_, lf = loop_freq_, ph_t / inc_t the tone LO phase and step, at a cost of
// 6 libm trig calls per sample.  Both phases advance linearly inside the symbol, so
//     corr_t = exp(-j(cp + ph_t)) * sum_i s[i] w_t^i,     w_t = exp(-j(lf + inc_t))
// is a 40-term Horner polynomial and two sincos per tone.  The soft decision Re(corr_2) - Re(corr_1)
// is phase-sensitive, so the phase accumulators are advanced exactly as the reference does it (40
// separate additions per symbol and the same wrap loops), the loop filter and AFC updates keep the
// reference's operation order and roundings, and the AFC's arg(dominant * conj(prev)) is evaluated
// with the reference's complex-product formula so that exact zeros keep their signs (all-zero
// input: atan2(+0, -0) = pi moves the offset by 27 Hz per symbol in the reference, too).
#pragma once
#include "opvd_common.cuh"

namespace opvd {

struct CoherentState {
    double freq_offset, carrier_phase, phase_f1, phase_f2, loop_freq;  // :549-553
    cplx prev;                                                         // prev_dominant_ :554
    double afc_alpha, pll_alpha, pll_beta;                             // :555-557
};

OPVD_HD void coherent_init(CoherentState& s, double freq_offset, double afc_alpha, double pll_bw_hz) {
    s.freq_offset = freq_offset; s.carrier_phase = 0.0; s.phase_f1 = 0.0; s.phase_f2 = 0.0; s.loop_freq = 0.0;
    s.prev = {0.0, 0.0};
    s.afc_alpha = afc_alpha;
    const double wn = pll_bw_hz * kTwoPi;  // set_pll_bandwidth, :561-568
    const double zeta = 0.707;
    s.pll_alpha = 2.0 * zeta * wn / kSymbolRate;
    s.pll_beta = wn * wn / (kSymbolRate * kSymbolRate);
}

OPVD_HD double coherent_wrap(double ph) {  // :486-491
    while (ph > kPi) ph -= kTwoPi;
    while (ph < -kPi) ph += kTwoPi;
    return ph;
}

// One symbol; I/Q: its 40 samples.  first_in_call: sym == 0 (no AFC update, :533).  Returns the soft symbol.
OPVD_HD double coherent_symbol(CoherentState& s, const double* I, const double* Q, bool first_in_call) {
    const double inc1 = kTwoPi * (-kFreqDev + s.freq_offset) / kSampleRate;  // :453-454, :539-540
    const double inc2 = kTwoPi * (+kFreqDev + s.freq_offset) / kSampleRate;
    cplx c[2];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        const double th = s.loop_freq + (t ? inc2 : inc1);
        double sn, cs;
        sincos(th, &sn, &cs);
        const cplx w = {cs, -sn};
        cplx g = {I[kSps - 1], Q[kSps - 1]};
        for (int i = kSps - 2; i >= 0; --i) {
            const double nr = fma(g.r, w.r, fma(-g.i, w.i, I[i]));
            const double ni = fma(g.r, w.i, fma(g.i, w.r, Q[i]));
            g.r = nr; g.i = ni;
        }
        if (g.r == 0.0 && g.i == 0.0) {
            c[t] = {0.0, 0.0};  // the reference's accumulators start at +0 and stay +0 on all-zero input
        } else {
            sincos(s.carrier_phase + (t ? s.phase_f2 : s.phase_f1), &sn, &cs);
            c[t] = cmul(cplx{cs, -sn}, g);
        }
    }
    // phase accumulators: 40 separate additions, as the reference's per-sample loop (:478-482)
    for (int i = 0; i < kSps; ++i) {
        s.phase_f1 += inc1;
        s.phase_f2 += inc2;
        s.carrier_phase += s.loop_freq;
    }
    s.phase_f1 = coherent_wrap(s.phase_f1);
    s.phase_f2 = coherent_wrap(s.phase_f2);
    s.carrier_phase = coherent_wrap(s.carrier_phase);

    const double e1 = c[0].r * c[0].r + c[0].i * c[0].i, e2 = c[1].r * c[1].r + c[1].i * c[1].i;  // :494-495
    const double soft = c[1].r - c[0].r;                                                           // :500-505
    const cplx dom = (e1 > e2) ? c[0] : c[1];                                                      // :510
    const double mag = sqrt(dom.r * dom.r + dom.i * dom.i);
    double phase_error = 0.0;
    if (mag > 1e-10) phase_error = dom.i / mag;                                                    // :514-520
    s.loop_freq += s.pll_beta * phase_error;                                                       // :524-525
    s.carrier_phase += s.pll_alpha * phase_error;
    s.loop_freq = clampd(s.loop_freq, -0.1, 0.1);                                                  // :528
    if (!first_in_call) {                                                                          // :533-541
        const double npi = -s.prev.i;  // dominant * conj(prev_dominant), std::complex operator*
        const double xr = dom.r * s.prev.r - dom.i * npi;
        const double xi = dom.r * npi + dom.i * s.prev.r;
        const double phase_diff = atan2(xi, xr);
        const double freq_err = phase_diff * kSymbolRate / kTwoPi;
        s.freq_offset += s.afc_alpha * freq_err;
        s.freq_offset = clampd(s.freq_offset, -2000.0, 2000.0);
    }
    s.prev = dom;                                                                                  // :543
    return soft;
}

}  // namespace opvd
