// demod_batch_core.cuh — arithmetic of the BATCHED demodulator (kernels_demod_batch.cu): a CTA owns 32
// streams; every symbol is split into a parallel window phase (four helper threads per stream: one
// tone x one half of the 60-sample window each) and a serial loop phase (one lane per stream).
// Same algorithm as demod_core.cuh (reference: MSKDemodulatorAFC::demodulate,
// /root/reference/src/opv-demod.cpp:206-329).  Host/device code: the CUDA kernel and the CPU
// host-sim test (tests/hostsim) compile exactly these functions.
//
// Window slots k = 0..60 are the raw samples b-10 .. b+50; slot k carries the weight z^k inside a
// 10-slot block and q = z^10 between blocks:
//   H_m = sum_{j<10} s[10m+j] z^j                       block sums, m = 0..5   (Horner, helper threads)
//   E = H0 + qH1 + q^2 H2 + q^3 H3,  O = H1 + .. + q^3 H4,  L = H2 + .. + q^3 H5
// A helper owns blocks 3h..3h+2 (h = window half) and hands over three partial gates so that the
// serial lane only needs three complex FMAs per tone:
//   h = 0:  Ea = H0 + q(H1 + qH2),  Oa = H1 + qH2,  La = H2
//   h = 1:  Eb = H3,  Ob = H3 + qH4,  Lb = H3 + q(H4 + qH5)
//   E = Ea + q^3 Eb,   O = Oa + q^2 Ob,   L = La + q Lb
// The post-sum linear interpolator X = C + f*(conj(z)*(C + dX) - C) (demod_core.cuh) is linear in the
// gate sum C and its shifted-window edge term dX = s[last+1] z^40 - s[first], and both split over the
// halves with the same weights (dE = -s0 + q^3 (s40 q), dO = -s10 + q^2 (s50 q^2), dL = -s20 + q (s60 q^3)),
// so each helper interpolates its own partial gates and the serial lane combines finished values.
#pragma once
#include "demod_core.cuh"
#include "fastmath.cuh"

namespace opvd {

struct HalfGates {
    cplx E, O, L;
};

// Horner over 10 consecutive samples
OPVD_HD cplx horner10(const double* I, const double* Q, cplx z) {
    cplx g = {I[9], Q[9]};
#pragma unroll
    for (int r = 8; r >= 0; --r) {
        const double nr = fma(g.r, z.r, fma(-g.i, z.i, I[r]));
        const double ni = fma(g.r, z.i, fma(g.i, z.r, Q[r]));
        g.r = nr; g.i = ni;
    }
    return g;
}

// The same block sum as two 5-sample Horner chains joined by z5 = z^5: same operation count (36), half the
// dependent depth — for kernels where one warp has to keep the FP64 pipe busy on its own.
OPVD_HD cplx horner10_split(const double* I, const double* Q, cplx z, cplx z5) {
    cplx lo = {I[4], Q[4]}, hi = {I[9], Q[9]};
#pragma unroll
    for (int r = 3; r >= 0; --r) {
        const double lr = fma(lo.r, z.r, fma(-lo.i, z.i, I[r]));
        const double li = fma(lo.r, z.i, fma(lo.i, z.r, Q[r]));
        const double hr = fma(hi.r, z.r, fma(-hi.i, z.i, I[r + 5]));
        const double hq = fma(hi.r, z.i, fma(hi.i, z.r, Q[r + 5]));
        lo.r = lr; lo.i = li; hi.r = hr; hi.i = hq;
    }
    return cfma(z5, hi, lo);
}

// one tone, one window half, from the half's three block sums A, B, C (slots 30h.., 30h+10.., 30h+20..) and
// the three raw samples its edge terms need: e0, e1, e2 = slots 0, 10, 20 (h = 0) or 40, 50, 60 (h = 1).
// Returns the interpolated partial gates.
OPVD_HD HalfGates half_gates_from_blocks(cplx A, cplx B, cplx C, cplx e0, cplx e1, cplx e2, cplx z, cplx q, double f,
                                         int half) {
    HalfGates g, d;  // partial gate sums and their edge terms
    if (half == 0) {
        g.L = C;
        g.O = cfma(q, C, B);
        g.E = cfma(q, g.O, A);
        d.E = {-e0.r, -e0.i};
        d.O = {-e1.r, -e1.i};
        d.L = {-e2.r, -e2.i};
    } else {
        g.E = A;
        g.O = cfma(q, B, A);
        g.L = cfma(q, cfma(q, C, B), A);
        const cplx q2 = csqr(q), q3 = cmul(q2, q);
        d.E = {q.r * e0.r, q.i * e0.r};
        d.E = {fma(-q.i, e0.i, d.E.r), fma(q.r, e0.i, d.E.i)};     // s40 * q
        d.O = {q2.r * e1.r, q2.i * e1.r};
        d.O = {fma(-q2.i, e1.i, d.O.r), fma(q2.r, e1.i, d.O.i)};   // s50 * q^2
        d.L = {q3.r * e2.r, q3.i * e2.r};
        d.L = {fma(-q3.i, e2.i, d.L.r), fma(q3.r, e2.i, d.L.i)};   // s60 * q^3
    }
    auto interp = [&](cplx Cg, cplx dX) {
        const cplx S = {Cg.r + dX.r, Cg.i + dX.i};
        const cplx T = {fma(z.r, S.r, z.i * S.i), fma(z.r, S.i, -(z.i * S.r))};  // conj(z) * S
        return cplx{fma(f, T.r - Cg.r, Cg.r), fma(f, T.i - Cg.i, Cg.i)};
    };
    HalfGates o;
    o.E = interp(g.E, d.E);
    o.O = interp(g.O, d.O);
    o.L = interp(g.L, d.L);
    return o;
}

// the same from the half's 30 samples I/Q (slots 30h .. 30h+29) plus, for h = 1, slot 60 in I[30]/Q[30]
OPVD_HD HalfGates batch_half_gates(const double* I, const double* Q, cplx z, cplx q, double f, int half) {
    const cplx A = horner10(I, Q, z), B = horner10(I + 10, Q + 10, z), C = horner10(I + 20, Q + 20, z);
    const cplx s0 = {I[0], Q[0]}, s10 = {I[10], Q[10]}, s20 = {I[20], Q[20]}, s30 = {I[30], Q[30]};
    return half_gates_from_blocks(A, B, C, half ? s10 : s0, half ? s20 : s10, half ? s30 : s20, z, q, f, half);
}

// tone step z = exp(-j*inc_t) and block step q = z^10 from the AFC offset (:210-211, :305-306)
struct ToneLo {
    cplx z, q;
    double inc;
    cplx z5;  // z^5 (horner10_split)
};
OPVD_HD void batch_lo(double freq_offset, ToneLo& t1, ToneLo& t2) {
    const LoSteps l = lo_steps(freq_offset);  // Taylor zeta, sincos fallback for huge -o offsets
    t1.z = l.z1; t2.z = l.z2; t1.inc = l.inc1; t2.inc = l.inc2;
    cplx a = csqr(l.z1), b = csqr(l.z2);      // z^2
    cplx a4 = csqr(a), b4 = csqr(b);          // z^4
    a = cmul(a4, l.z1); b = cmul(b4, l.z2);   // z^5
    t1.z5 = a; t2.z5 = b;
    t1.q = csqr(a); t2.q = csqr(b);           // z^10
}

// LO steps in the hot loop: valid for |freq_offset| <= 2.5 kHz, which the AFC clamp (:303) guarantees
// after every update; coefficients come from the constant table (no FP64 immediates in the loop).
OPVD_HD void batch_lo_fast(double freq_offset, ToneLo& t1, ToneLo& t2, const FastMathTable& K) {
    const double d = freq_offset * K.two_pi_over_fs;
    const cplx zeta = expmj_small(d, K);
    t1.inc = d - K.inc_dev; t2.inc = d + K.inc_dev;
    t1.z = cmul(cplx{K.tau_c, K.tau_s}, zeta);    // exp(+j*2pi/160) * exp(-j*delta)
    t2.z = cmul(cplx{K.tau_c, -K.tau_s}, zeta);
    cplx a = csqr(t1.z), b = csqr(t2.z);      // z^2
    const cplx a4 = csqr(a), b4 = csqr(b);    // z^4
    a = cmul(a4, t1.z); b = cmul(b4, t2.z);   // z^5
    t1.z5 = a; t2.z5 = b;
    t1.q = csqr(a); t2.q = csqr(b);           // z^10
}

// Serial lane: combine the two halves' interpolated partial gates of one tone; returns the gate
// energies, the on-time sum and z^40 (for the AFC's previous-correlation rotation).
struct ToneGates {
    cplx O;          // interpolated on-time correlation (common unit-modulus phase factor dropped)
    double eE, eO, eL;
    cplx z40;
};
OPVD_HD ToneGates batch_finish_tone(const HalfGates& a, const HalfGates& b, const ToneLo& t, cplx fixE) {
    const cplx q2 = csqr(t.q);
    const cplx q3 = cmul(q2, t.q);
    ToneGates o;
    o.z40 = csqr(q2);
    cplx E = cfma(q3, b.E, a.E);
    const cplx L = cfma(t.q, b.L, a.L);
    o.O = cfma(q2, b.O, a.O);
    E.r -= fixE.r; E.i -= fixE.i;  // early-gate clamp of the first symbol of a call (:237); zero otherwise
    o.eE = cnorm(E); o.eO = cnorm(o.O); o.eL = cnorm(L);
    return o;
}

// Loop-carried registers of one stream in the serial lane.
struct BatchRegs {
    double freq_offset, ph1, ph2, pos, timing_freq;
    cplx p1, p2;
    ToneLo t1, t2;
};

OPVD_HD double clamp_sym_b(double v, double lim) { return fabs(v) > lim ? copysign(lim, v) : v; }

OPVD_HD_COLD double batch_afc_corner(cplx dom, cplx prev, double ph) { return afc_phase_signed_zero(dom, prev, ph); }

// The serial part of one symbol (:264-313) splits into two independent chains that the kernel runs on
// two warps: the timing chain (soft decision, TED, timing loop -> pos) and the AFC chain (phase
// detector, AFC loop, previous correlations, LO phases -> freq_offset).

// timing chain (:264-286, :313); returns the soft symbol
OPVD_HD double batch_timing(double eO1, double eO2, double eE1, double eL1, double eE2, double eL2,
                            double& timing_freq, double& pos, const FastMathTable& K) {
    const bool tone1 = eO1 > eO2;          // :272
    const double ee = tone1 ? eE1 : eE2, el = tone1 ? eL1 : eL2;
    const double ted = div_fast(el - ee, el + ee + K.eps_ted);           // :280
    timing_freq = clamp_sym_b(timing_freq + K.k_tf * ted, K.lim_tf);     // :283-284
    const double adj = clamp_sym_b(K.k_adj * ted + timing_freq, 2.0);    // :285-286
    pos += 40.0 + adj;                                                   // :313
    return eO2 - eO1;                                                    // :268
}

// AFC chain (:289-310, :250-262).  O1/O2: interpolated on-time sums, z40_t = z_t^40, inc_t = LO phase
// steps used during this symbol.  first_in_call: no AFC update (:289).
struct BatchAfc {
    double freq_offset, ph1, ph2;
    cplx p1, p2;
};
OPVD_HD void batch_afc(BatchAfc& r, cplx O1, cplx z40_1, double eO1, cplx O2, cplx z40_2, double eO2, double inc1,
                       double inc2, bool first_in_call, double afc_alpha, const FastMathTable& K) {
    if (!first_in_call) {                                               // :289-307
        const bool tone1 = eO1 > eO2;                                   // :291
        const cplx dom = tone1 ? O1 : O2, prev = tone1 ? r.p1 : r.p2;
        const double xr = fma(dom.r, prev.r, dom.i * prev.i);
        const double xi = fma(dom.i, prev.r, -(dom.r * prev.i));
        double pd = atan2_fast(xi, xr, K);
        const bool corner = (dom.r == 0.0 && dom.i == 0.0) || (prev.r == 0.0 && prev.i == 0.0);
        if (corner) pd = batch_afc_corner(dom, prev, tone1 ? r.ph1 : r.ph2);
        const double ferr = pd * K.sym_rate_over_two_pi;
        r.freq_offset = clamp_sym_b(r.freq_offset + afc_alpha * ferr, 2000.0);
    }
    // previous correlations for the NEXT symbol, rotated to the phase frame at the next symbol start
    r.p1 = cmul(O1, cconj(z40_1));  // :309-310
    r.p2 = cmul(O2, cconj(z40_2));
    const double a1 = fma(40.0, inc1, r.ph1), a2 = fma(40.0, inc2, r.ph2);  // :250-262
    r.ph1 = fma(-K.two_pi, rint(a1 * K.inv_two_pi), a1);
    r.ph2 = fma(-K.two_pi, rint(a2 * K.inv_two_pi), a2);
}

// Both chains on one lane (host simulation and single-warp use).
OPVD_HD double batch_symbol_serial(BatchRegs& r, const ToneGates& g1, const ToneGates& g2, bool first_in_call,
                                   double afc_alpha, const FastMathTable& K) {
    const double soft = batch_timing(g1.eO, g2.eO, g1.eE, g1.eL, g2.eE, g2.eL, r.timing_freq, r.pos, K);
    BatchAfc a = {r.freq_offset, r.ph1, r.ph2, r.p1, r.p2};
    batch_afc(a, g1.O, g1.z40, g1.eO, g2.O, g2.z40, g2.eO, r.t1.inc, r.t2.inc, first_in_call, afc_alpha, K);
    r.freq_offset = a.freq_offset; r.ph1 = a.ph1; r.ph2 = a.ph2; r.p1 = a.p1; r.p2 = a.p2;
    if (!first_in_call) batch_lo_fast(r.freq_offset, r.t1, r.t2, K);  // |freq_offset| <= 2 kHz after the clamp
    return soft;
}

}  // namespace opvd
