// demod_core.cuh — per-symbol arithmetic of the dual-tone MSK demodulator with AFC and
// early-late symbol timing recovery (reference: MSKDemodulatorAFC::demodulate,
// /root/reference/src/opv-demod.cpp:206-329), restructured for one-stream-per-lane execution
// on B200's FP64 pipe.  Host/device code: the CUDA kernels and the CPU host-sim test compile
// exactly this arithmetic.
//
// What the reference computes per symbol (position pos = b + f, b integer, 0 <= f < 1):
//     corr_t   = sum_{i<40} y[i]    * exp(-j(ph_t + i*inc_t))        on-time, tone t in {1,2}
//     early_t  = sum_{i<40} y[i-10] * exp(-j(ph_t + i*inc_t))
//     late_t   = sum_{i<40} y[i+10] * exp(-j(ph_t + i*inc_t))
//     y[k]     = (1-f)*s[b+k] + f*s[b+k+1]                           linear interpolation (:122-128)
// at a cost of 3 interpolations, 4 libm trig calls and 6 complex MACs per sample.  Only
// |corr|^2, |early|^2, |late|^2 and arg(corr_n * conj(corr_{n-1})) are ever used (:264-306).
//
// Restructuring (exact in real arithmetic, ~1e-15 relative in FP64):
//  * z = exp(-j*inc_t).  All three gates are polynomials in z over the SAME 61 raw samples
//    s[b-10 .. b+50]; the common factor exp(-j*ph_t) has modulus 1 and drops out of the norms.
//  * Each polynomial is evaluated by Horner over 12 segments of 5 samples (4 FMAs per sample and
//    tone, no per-sample trig), the segments are combined with z^5, z^10, z^20.
//  * The interpolator is linear, so it is applied AFTER the sums:
//        sum_k y[k] z^k = g*X + h*dX,  g = (1-f) + f*conj(z),  h = f*conj(z),
//        dX = -s[first] + s[last+1]*z^40
//  * AFC needs corr_n*conj(corr_{n-1}); the phase advance between two symbol starts is 40*inc_t
//    of the earlier symbol, so the previous on-time sum is stored pre-rotated by conj(z^40).
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
    int32_t flags;        // kFlagDone | kFlagEstDone
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

// e^{-j*inc} for both tones from the current AFC offset (:210-211, :305-306)
struct LoSteps {
    cplx z1, z2;
    double inc1, inc2;
};

OPVD_HD LoSteps lo_steps(double freq_offset) {
    LoSteps l;
    l.inc1 = kTwoPi * (-kFreqDev + freq_offset) / kSampleRate;
    l.inc2 = kTwoPi * (+kFreqDev + freq_offset) / kSampleRate;
    double s, c;
    sincos(l.inc1, &s, &c);
    l.z1 = {c, -s};
    sincos(l.inc2, &s, &c);
    l.z2 = {c, -s};
    return l;
}

OPVD_HD double wrap_phase(double ph) {  // :259-262
    while (ph > kPi) ph -= kTwoPi;
    while (ph < -kPi) ph += kTwoPi;
    return ph;
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
    // shifted-window corrections dX = s[last+1]*z^40 - s[first]
    cplx dE = {fma(sI[3], p.z40.r, fma(-sQ[3], p.z40.i, -sI[0])), fma(sI[3], p.z40.i, fma(sQ[3], p.z40.r, -sQ[0]))};
    cplx dO = {fma(sI[4], p.z40.r, fma(-sQ[4], p.z40.i, -sI[1])), fma(sI[4], p.z40.i, fma(sQ[4], p.z40.r, -sQ[1]))};
    cplx dL = {fma(sI[5], p.z40.r, fma(-sQ[5], p.z40.i, -sI[2])), fma(sI[5], p.z40.i, fma(sQ[5], p.z40.r, -sQ[2]))};
    cplx h = {f * p.z.r, -(f * p.z.i)};   // f*conj(z)
    cplx g = {(1.0 - f) + h.r, h.i};      // (1-f) + f*conj(z)
    Gates o;
    o.E = cfma(g, E, cmul(h, dE));
    o.O = cfma(g, O, cmul(h, dO));
    o.L = cfma(g, L, cmul(h, dL));
    return o;
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

// Loop-carried registers of one stream while a kernel is running.
struct DemodRegs {
    double freq_offset, ph1, ph2, pos, timing_freq;
    cplx p1, p2;
    LoSteps lo;
};

// One symbol.  win[0..60] = packed raw samples at local indices b-10 .. b+50 (b = floor(pos)).
// first_in_call: the early gate must see samples[0] for local indices < 0 (:237) and the AFC update
// is skipped (:289); `s0` is the packed sample at local index 0 in that case.
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
        // early gate: y[k] := samples[0] for k = -10..-1  (:237).  Remove what the generic path summed
        // for those ten taps and add the clamped value instead.
        double I0 = sI[1], Q0 = sQ[1];
        cplx fix1 = {0.0, 0.0}, fix2 = {0.0, 0.0};
        for (int k = 9; k >= 0; --k) {  // Horner over exponents k = 0..9 (tap local index k-10)
            double Ia, Qa, Ib, Qb;
            unpack_iq(win[k], Ia, Qa);
            unpack_iq(win[k + 1], Ib, Qb);
            double yr = fma(f, Ib - Ia, Ia) - I0;  // y[k-10] - s0
            double yi = fma(f, Qb - Qa, Qa) - Q0;
            fix1 = {fma(fix1.r, pw1.z.r, fma(-fix1.i, pw1.z.i, yr)), fma(fix1.r, pw1.z.i, fma(fix1.i, pw1.z.r, yi))};
            fix2 = {fma(fix2.r, pw2.z.r, fma(-fix2.i, pw2.z.i, yr)), fma(fix2.r, pw2.z.i, fma(fix2.i, pw2.z.r, yi))};
        }
        g1.E.r -= fix1.r; g1.E.i -= fix1.i;
        g2.E.r -= fix2.r; g2.E.i -= fix2.i;
    }

    const double e1 = cnorm(g1.O);
    const double e2 = cnorm(g2.O);
    const double soft = e2 - e1;  // :268
    const bool tone1 = e1 > e2;   // :272, :291

    // timing error detector + 2nd-order loop (:271-286)
    const cplx ge = tone1 ? g1.E : g2.E;
    const cplx gl = tone1 ? g1.L : g2.L;
    const double ee = cnorm(ge), el = cnorm(gl);
    const double ted = (el - ee) / (el + ee + 1e-10);
    r.timing_freq += 0.00001 * ted;
    r.timing_freq = clampd(r.timing_freq, -0.1, 0.1);
    double timing_adj = 0.005 * ted + r.timing_freq;
    timing_adj = clampd(timing_adj, -2.0, 2.0);

    // previous correlations for the NEXT symbol: rotate this symbol's on-time sums to the phase
    // frame at the next symbol start (phase advances by 40*inc of THIS symbol)
    const cplx n1 = cmul(g1.O, cconj(pw1.z40));
    const cplx n2 = cmul(g2.O, cconj(pw2.z40));

    if (!first_in_call) {  // :289-307
        const cplx dom = tone1 ? g1.O : g2.O;
        const cplx prev = tone1 ? r.p1 : r.p2;
        double pd;
        const bool dz = (dom.r == 0.0 && dom.i == 0.0), pz = (prev.r == 0.0 && prev.i == 0.0);
        if (dz || pz) {
            pd = afc_phase_signed_zero(dom, prev, tone1 ? r.ph1 : r.ph2);
        } else {
            const double xr = fma(dom.r, prev.r, dom.i * prev.i);
            const double xi = fma(dom.i, prev.r, -(dom.r * prev.i));
            pd = atan2(xi, xr);
        }
        const double ferr = pd * kSymbolRate / kTwoPi;
        r.freq_offset += afc_alpha * ferr;
        r.freq_offset = clampd(r.freq_offset, -2000.0, 2000.0);
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
                s.call_len = avail;        // may be 0: the loop condition below closes it at once
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
