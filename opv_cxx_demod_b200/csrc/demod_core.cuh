// demod_core.cuh — per-symbol arithmetic of the dual-tone MSK demodulator with AFC and
// early-late symbol timing recovery (reference: MSKDemodulatorAFC::demodulate,
// /root/reference/src/opv-demod.cpp:206-329), restructured for B200's FP64 pipe.
// Host/device code: the CUDA kernels and the CPU host-sim test compile exactly this arithmetic.
//
// What the reference computes per symbol (position pos = b + f, b integer, 0 <= f < 1):
//     corr_t   = sum_{i<40} y[i]    * exp(-j(ph_t + i*inc_t))        on-time, tone t in {1,2}
//     early_t  = sum_{i<40} y[i-10] * exp(-j(ph_t + i*inc_t))
//     late_t   = sum_{i<40} y[i+10] * exp(-j(ph_t + i*inc_t))
//     y[k]     = (1-f)*s[b+k] + f*s[b+k+1]                           linear interpolation (:122-128)
// at a cost of 3 interpolations, 4 libm trig calls and 6 complex MACs per sample.  Only
// |corr|^2, |early|^2, |late|^2 and arg(corr_n * conj(corr_{n-1})) are ever used (:264-306).
//
// Restructuring (exact in real arithmetic, ~1e-12 of rms in FP64 after the loops have fed back):
//  * z = exp(-j*inc_t).  All three gates are polynomials in z over the SAME 61 raw samples
//    s[b-10 .. b+50]; the common factor exp(-j*ph_t) has modulus 1 and drops out of the norms.
//  * Each polynomial is evaluated by Horner over 12 segments of 5 samples (4 FMAs per sample and
//    tone, no per-sample trig), the segments are combined with z^5, z^10, z^20.
//  * The interpolator is linear, so it is applied AFTER the sums:
//        sum_k y[k] z^k = g*X + h*dX,  g = (1-f) + f*conj(z),  h = f*conj(z),
//        dX = -s[first] + s[last+1]*z^40
//  * AFC needs corr_n*conj(corr_{n-1}); the phase advance between two symbol starts is 40*inc_t
//    of the earlier symbol, so the previous on-time sum is stored pre-rotated by conj(z^40).
//  * The tone steps are exactly -/+ 2*pi/160 plus the AFC term delta = 2*pi*offset/fs with
//    |delta| <= 0.006, so z_t = tau^(+/-1) * zeta with tau a constant and zeta = exp(-j*delta)
//    from a short Taylor polynomial: no sincos call on the per-symbol critical path.
//  * The absolute LO phases are still tracked (2 FMAs per symbol) because the reference's
//    atan2 sees signed zeros when a correlation is exactly 0 (all-zero input, e.g. the 4000
//    trailing zeros opv-mod appends): the +/-pi it then returns depends on the quadrant of the
//    absolute-phase correlation.  That corner is reproduced in afc_phase_signed_zero().
#pragma once
#include "opvd_common.cuh"

namespace opvd {

// Persistent per-stream demodulator state (what MSKDemodulatorAFC keeps between calls, :337-347,
// plus the streaming driver's chunk bookkeeping, :1012-1076).
struct DemodState {
    double freq_offset;   // :337
    double ph1, ph2;      // :338 absolute LO phases at the next symbol start
    cplx p1, p2;          // :339 previous on-time correlations, in the next symbol's phase frame
    double pos;           // position inside the open call; == mu_ (:343) between calls
    double timing_freq;   // :344
    int64_t origin;       // global sample index of sample 0 of the open / next call
    int64_t call_len;     // N of the open call (0 = none open)
    int64_t n_sym;        // soft symbols produced so far (global symbol index of the next one)
    int32_t sym_in_call;  // symbols produced in the open call; 0 => AFC update skipped (:289)
    int32_t flags;        // kFlagDone | kFlagEstDone | kFlagFlush
};
constexpr int32_t kFlagDone = 1;     // EOF flush performed (stream mode) / single call finished (batch)
constexpr int32_t kFlagEstDone = 2;  // initial offset decided (estimate, -o, or "never" for short streams)
constexpr int32_t kFlagFlush = 4;    // the open call is the EOF flush (:1088-1113)

OPVD_HD void demod_state_init(DemodState& s) {
    s.freq_offset = 0.0; s.ph1 = 0.0; s.ph2 = 0.0;
    s.p1 = {0.0, 0.0}; s.p2 = {0.0, 0.0};
    s.pos = 0.0; s.timing_freq = 0.0;
    s.origin = 0; s.call_len = 0; s.n_sym = 0; s.sym_in_call = 0; s.flags = 0;
}

// ---------------------------------------------------------------------------------------------
// LO steps.  inc_t = 2*pi*(-/+13550 + offset)/fs (:210-211, :305-306); 13550/fs = 1/160 exactly.
constexpr double kTauC = 0.9992290362407229;         // cos(2*pi/160)
constexpr double kTauS = 0.03925981575906861;        // sin(2*pi/160)
constexpr double kIncDev = 0.039269908169872414;     // 2*pi/160
constexpr double kTwoPiOverFs = 2.8981482044186284e-06;
constexpr double kSymRateOverTwoPi = 8626.197915580728;  // 54200 / (2*pi)  (:300)
constexpr double kInvTwoPi = 0.15915494309189535;

// zeta = exp(-j*delta), delta = 2*pi*offset/fs.  Taylor to delta^12: exact to < 1e-17 for
// |offset| <= 50 kHz (the AFC clamps to +/-2 kHz, :303; a larger -o value is possible before the
// first update), sincos() beyond that.
OPVD_HD cplx zeta_from_offset(double offset_hz, double& delta) {
    const double d = offset_hz * kTwoPiOverFs;
    delta = d;
    if (fabs(d) > 0.145) {
        double s, c;
        sincos(d, &s, &c);
        return {c, -s};
    }
    const double d2 = d * d;
    double s = fma(d2, -1.0 / 39916800.0, 1.0 / 362880.0);
    s = fma(d2, s, -1.0 / 5040.0);
    s = fma(d2, s, 1.0 / 120.0);
    s = fma(d2, s, -1.0 / 6.0);
    s = fma(d2 * d, s, d);
    double c = fma(d2, 1.0 / 479001600.0, -1.0 / 3628800.0);
    c = fma(d2, c, 1.0 / 40320.0);
    c = fma(d2, c, -1.0 / 720.0);
    c = fma(d2, c, 1.0 / 24.0);
    c = fma(d2, c, -0.5);
    c = fma(d2, c, 1.0);
    return {c, -s};
}

struct LoSteps {
    cplx z1, z2;        // exp(-j*inc1), exp(-j*inc2)
    double inc1, inc2;
};

OPVD_HD LoSteps lo_steps(double freq_offset) {
    LoSteps l;
    double d;
    const cplx zeta = zeta_from_offset(freq_offset, d);
    l.inc1 = d - kIncDev;
    l.inc2 = d + kIncDev;
    l.z1 = cmul(cplx{kTauC, kTauS}, zeta);   // exp(+j*2pi/160) * exp(-j*delta)
    l.z2 = cmul(cplx{kTauC, -kTauS}, zeta);
    return l;
}

// one tone only (tone 0 -> F1, tone 1 -> F2): used when the two tones live on different lanes
OPVD_HD void lo_step_tone(double freq_offset, int tone, cplx& z, double& inc) {
    double d;
    const cplx zeta = zeta_from_offset(freq_offset, d);
    inc = tone ? d + kIncDev : d - kIncDev;
    z = cmul(cplx{kTauC, tone ? -kTauS : kTauS}, zeta);
}

OPVD_HD double wrap_phase(double ph) {  // (-pi, pi] up to rounding; only the signed-zero corner reads it
    return fma(-kTwoPi, rint(ph * kInvTwoPi), ph);
}

// Horner over 5 consecutive samples: s0 + z*(s1 + z*(s2 + z*(s3 + z*s4)))
OPVD_HD cplx horner5(const double* I, const double* Q, cplx z) {
    cplx g = {I[4], Q[4]};
#pragma unroll
    for (int r = 3; r >= 0; --r) {
        double nr = fma(g.r, z.r, fma(-g.i, z.i, I[r]));
        double ni = fma(g.r, z.i, fma(g.i, z.r, Q[r]));
        g.r = nr; g.i = ni;
    }
    return g;
}

struct TonePowers {
    cplx z, w5, q, q2, z40;
};

OPVD_HD TonePowers tone_powers(cplx z) {
    TonePowers p;
    p.z = z;
    cplx z2 = csqr(z), z4 = csqr(z2);
    p.w5 = cmul(z4, z);
    p.q = csqr(p.w5);     // z^10
    p.q2 = csqr(p.q);     // z^20
    p.z40 = csqr(p.q2);   // z^40
    return p;
}

struct Gates {
    cplx E, O, L;  // interpolated early / on-time / late sums (common unit-modulus phase factor dropped)
};

// s*z40 - f0  for real-pair samples: shifted-window correction dX = s[last+1]*z^40 - s[first]
OPVD_HD cplx edge_term(double lastI, double lastQ, double firstI, double firstQ, cplx z40) {
    return {fma(lastI, z40.r, fma(-lastQ, z40.i, -firstI)), fma(lastI, z40.i, fma(lastQ, z40.r, -firstQ))};
}

// interpolator applied after the sums: g*X + h*dX
OPVD_HD void interp_weights(cplx z, double f, cplx& g, cplx& h) {
    h = {f * z.r, -(f * z.i)};       // f*conj(z)
    g = {(1.0 - f) + h.r, h.i};      // (1-f) + f*conj(z)
}

// combine six 10-sample partial sums H[m] (samples 10m-10 .. 10m-1 relative to b, exponent origin at
// the segment start) into the three 40-sample gates and apply the interpolator.
// sI/sQ: raw samples at local indices -10, 0, 10, 30, 40, 50  (window slots 0,10,20,40,50,60)
OPVD_HD Gates combine_gates(const cplx* H, const TonePowers& p, double f, const double* sI, const double* sQ) {
    cplx T01 = cfma(p.q, H[1], H[0]);
    cplx T12 = cfma(p.q, H[2], H[1]);
    cplx T23 = cfma(p.q, H[3], H[2]);
    cplx T34 = cfma(p.q, H[4], H[3]);
    cplx T45 = cfma(p.q, H[5], H[4]);
    cplx E = cfma(p.q2, T23, T01);
    cplx O = cfma(p.q2, T34, T12);
    cplx L = cfma(p.q2, T45, T23);
    cplx dE = edge_term(sI[3], sQ[3], sI[0], sQ[0], p.z40);
    cplx dO = edge_term(sI[4], sQ[4], sI[1], sQ[1], p.z40);
    cplx dL = edge_term(sI[5], sQ[5], sI[2], sQ[2], p.z40);
    cplx g, h;
    interp_weights(p.z, f, g, h);
    Gates o;
    o.E = cfma(g, E, cmul(h, dE));
    o.O = cfma(g, O, cmul(h, dO));
    o.L = cfma(g, L, cmul(h, dL));
    return o;
}

// early-gate correction for the first symbol of a call: y[k] := samples[0] for local k = -10..-1 (:237).
// Returns sum_{k=0..9} (y[k-10] - s0) z^k, to be subtracted from the generic early sum.
OPVD_HD cplx first_symbol_fix(const uint32_t* win, double f, cplx z) {
    double I0, Q0;
    unpack_iq(win[kWinLead], I0, Q0);
    cplx fix = {0.0, 0.0};
    for (int k = 9; k >= 0; --k) {
        double Ia, Qa, Ib, Qb;
        unpack_iq(win[k], Ia, Qa);
        unpack_iq(win[k + 1], Ib, Qb);
        const double yr = fma(f, Ib - Ia, Ia) - I0;
        const double yi = fma(f, Qb - Qa, Qa) - Q0;
        fix = {fma(fix.r, z.r, fma(-fix.i, z.i, yr)), fma(fix.r, z.i, fma(fix.i, z.r, yi))};
    }
    return fix;
}

// timing error detector on the dominant tone's gates (:271-280)
OPVD_HD double ted_from_gates(cplx early, cplx late) {
    const double ee = cnorm(early), el = cnorm(late);
    return (el - ee) / (el + ee + 1e-10);
}

// AFC phase detector when dom or prev is exactly zero: reproduce the reference's signed zeros.
// ref corr = exp(-j*ph) * (our sum); an exactly-zero reference correlation is (+0,+0) because its
// accumulators start at +0 and +0 + (+/-0) = +0 (:223, :243).
OPVD_HD double afc_phase_signed_zero(cplx dom, cplx prev, double ph) {
    double s, c;
    sincos(ph, &s, &c);
    cplx rot = {c, -s};
    cplx d = {0.0, 0.0}, p = {0.0, 0.0};
    if (!(dom.r == 0.0 && dom.i == 0.0)) d = cmul(rot, dom);
    if (!(prev.r == 0.0 && prev.i == 0.0)) p = cmul(rot, prev);
    // dom * conj(prev) exactly as the compiler evaluates std::complex operator* (:299)
    double npi = -p.i;
    double xr = d.r * p.r - d.i * npi;
    double xi = d.r * npi + d.i * p.r;
    return atan2(xi, xr);
}

// arg(dom * conj(prev)) (:299)
OPVD_HD double afc_phase(cplx dom, cplx prev, double ph) {
    const bool dz = (dom.r == 0.0 && dom.i == 0.0), pz = (prev.r == 0.0 && prev.i == 0.0);
    if (dz || pz) return afc_phase_signed_zero(dom, prev, ph);
    const double xr = fma(dom.r, prev.r, dom.i * prev.i);
    const double xi = fma(dom.i, prev.r, -(dom.r * prev.i));
    return atan2(xi, xr);
}

// 2nd-order timing loop (:283-286); returns timing_adj
OPVD_HD double timing_loop(double& timing_freq, double ted) {
    timing_freq += 0.00001 * ted;
    timing_freq = clampd(timing_freq, -0.1, 0.1);
    return clampd(0.005 * ted + timing_freq, -2.0, 2.0);
}

// AFC loop (:300-303)
OPVD_HD void afc_loop(double& freq_offset, double pd, double afc_alpha) {
    const double ferr = pd * kSymRateOverTwoPi;
    freq_offset += afc_alpha * ferr;
    freq_offset = clampd(freq_offset, -2000.0, 2000.0);
}

// Loop-carried registers of one stream while a kernel is running (both tones on one lane).
struct DemodRegs {
    double freq_offset, ph1, ph2, pos, timing_freq;
    cplx p1, p2;
    LoSteps lo;
};

// One symbol, both tones on this lane.  win[0..60] = packed raw samples at local indices b-10 .. b+50.
// first_in_call: early-gate clamp (:237) and no AFC update (:289).
OPVD_HD double demod_symbol(DemodRegs& r, const uint32_t* win, double f, bool first_in_call, double afc_alpha) {
    const TonePowers pw1 = tone_powers(r.lo.z1);
    const TonePowers pw2 = tone_powers(r.lo.z2);

    cplx H1[6], H2[6];
    double sI[6], sQ[6];
#pragma unroll
    for (int m = 0; m < 6; ++m) {
        double I[10], Q[10];
#pragma unroll
        for (int k = 0; k < 10; ++k) unpack_iq(win[10 * m + k], I[k], Q[k]);
        if (m == 0) { sI[0] = I[0]; sQ[0] = Q[0]; }
        if (m == 1) { sI[1] = I[0]; sQ[1] = Q[0]; }
        if (m == 2) { sI[2] = I[0]; sQ[2] = Q[0]; }
        if (m == 4) { sI[3] = I[0]; sQ[3] = Q[0]; }
        if (m == 5) { sI[4] = I[0]; sQ[4] = Q[0]; }
        cplx a1 = horner5(I, Q, pw1.z), b1 = horner5(I + 5, Q + 5, pw1.z);
        cplx a2 = horner5(I, Q, pw2.z), b2 = horner5(I + 5, Q + 5, pw2.z);
        H1[m] = cfma(pw1.w5, b1, a1);
        H2[m] = cfma(pw2.w5, b2, a2);
    }
    unpack_iq(win[60], sI[5], sQ[5]);

    Gates g1 = combine_gates(H1, pw1, f, sI, sQ);
    Gates g2 = combine_gates(H2, pw2, f, sI, sQ);

    if (first_in_call) {
        const cplx fix1 = first_symbol_fix(win, f, pw1.z), fix2 = first_symbol_fix(win, f, pw2.z);
        g1.E.r -= fix1.r; g1.E.i -= fix1.i;
        g2.E.r -= fix2.r; g2.E.i -= fix2.i;
    }

    const double e1 = cnorm(g1.O);
    const double e2 = cnorm(g2.O);
    const double soft = e2 - e1;  // :268
    const bool tone1 = e1 > e2;   // :272, :291

    const double ted = ted_from_gates(tone1 ? g1.E : g2.E, tone1 ? g1.L : g2.L);
    const double timing_adj = timing_loop(r.timing_freq, ted);

    // previous correlations for the NEXT symbol: rotate this symbol's on-time sums to the phase
    // frame at the next symbol start (phase advances by 40*inc of THIS symbol)
    const cplx n1 = cmul(g1.O, cconj(pw1.z40));
    const cplx n2 = cmul(g2.O, cconj(pw2.z40));

    if (!first_in_call) {  // :289-307
        const double pd = afc_phase(tone1 ? g1.O : g2.O, tone1 ? r.p1 : r.p2, tone1 ? r.ph1 : r.ph2);
        afc_loop(r.freq_offset, pd, afc_alpha);
    }
    // absolute phases advance with the increments used during this symbol (:250-262)
    r.ph1 = wrap_phase(fma(40.0, r.lo.inc1, r.ph1));
    r.ph2 = wrap_phase(fma(40.0, r.lo.inc2, r.ph2));
    if (!first_in_call) r.lo = lo_steps(r.freq_offset);
    r.p1 = n1;  // :309-310
    r.p2 = n2;
    r.pos += 40.0 + timing_adj;  // :313
    return soft;
}

OPVD_HD void regs_from_state(DemodRegs& r, const DemodState& s) {
    r.freq_offset = s.freq_offset; r.ph1 = s.ph1; r.ph2 = s.ph2;
    r.pos = s.pos; r.timing_freq = s.timing_freq; r.p1 = s.p1; r.p2 = s.p2;
    r.lo = lo_steps(s.freq_offset);
}

OPVD_HD void regs_to_state(const DemodRegs& r, DemodState& s) {
    s.freq_offset = r.freq_offset; s.ph1 = r.ph1; s.ph2 = r.ph2;
    s.pos = r.pos; s.timing_freq = r.timing_freq; s.p1 = r.p1; s.p2 = r.p2;
}

// ---------------------------------------------------------------------------------------------
// Call scheduling: which demodulate() call a stream is in and when it ends.
//   batch  (:1164-1173): one call over the whole capture, N = total samples.
//   stream (:1012-1113): calls of exactly 86,720 samples starting at `origin`; after each call
//     origin += floor(pos) and pos = frac(pos) (the `leftover` carry, :318-328, :1070-1076);
//     at EOF one more call over whatever remains (:1088-1113).
// Returns true when a symbol can be demodulated now; false when the stream must wait for more
// samples or is finished (kFlagDone set).
OPVD_HD bool demod_schedule(DemodState& s, double& pos, int mode, int64_t avail, bool final) {
    for (;;) {
        if (s.flags & kFlagDone) return false;
        if (s.call_len == 0) {  // open the next call
            const int64_t remaining = avail - s.origin;
            if (mode == kModeBatch) {
                if (!final) return false;  // batch = load everything, then process
                s.call_len = avail;
                if (avail == 0) { s.flags |= kFlagDone; return false; }
            } else if (remaining >= kChunkSamples) {
                s.call_len = kChunkSamples;
            } else if (final && remaining > 0) {
                s.call_len = remaining;
                s.flags |= kFlagFlush;
            } else {
                if (final) s.flags |= kFlagDone;
                return false;
            }
            s.sym_in_call = 0;
        }
        if ((pos + 40.0) + 10.0 < (double)s.call_len) return true;  // :221
        // close the call (:318-328)
        const int64_t used = (int64_t)pos;
        pos = pos - (double)used;
        const bool was_flush = (s.flags & kFlagFlush) != 0;
        s.call_len = 0;
        if (mode == kModeBatch || was_flush) { s.flags |= kFlagDone; return false; }
        s.origin += used;  // leftover = N - used samples stay at the head of the next chunk
    }
}

}  // namespace opvd
