// demod_warp_core.cuh — per-lane arithmetic of the WARP-PER-STREAM demodulator (low per-symbol
// latency; used when there are too few streams to fill the machine with one lane per stream).
// Same algorithm and algebra as demod_core.cuh (reference: MSKDemodulatorAFC::demodulate,
// /root/reference/src/opv-demod.cpp:206-329), split across the lanes of one warp.
// Host/device code: the CUDA kernel (kernels_demod_warp.cu) and the CPU host-sim test
// (tests/hostsim) compile exactly these functions; only the lane exchange differs (warp shuffles
// on the device, array indexing on the host).
//
// Lane map: tone t = lane >> 4 (0 -> F1, 1 -> F2), p = lane & 15.
//   p = 0..11  own window slots c..c+4, c = 5p  (window slot k = raw sample b-10+k, k = 0..60)
//   p = 12     owns slot 60 only (the interpolator's extra sample of the late gate)
//   p = 13..15 idle
// Every slot k carries the weight z^k (z = exp(-j*inc_t)):
//   W_p = R_c * Horner5(s[c..c+4], z),   R_c = z^c        lane partial sum
//   F_p = R_c * s[c]                                      first-sample term (edge corrections)
// The three gates are sums of 8 consecutive lane partials,
//   E' = W_0+..+W_7 (slots 0..39), O' = W_2+..+W_9 (10..49), L' = W_4+..+W_11 (20..59),
// formed by three shuffle-down steps (1, 2, 4) and landing on lanes p = 0, 2, 4.  With the
// shifted-window term dX' = F_{p+8} - F_p the interpolated gate is X = g*X' + h*dX'
// (g = (1-f) + f*conj(z), h = f*conj(z)), identical to demod_core.cuh up to the unit-modulus
// factor z^c of the gate's first slot, which only matters for the AFC phase detector and is
// removed there.  R_c is NOT built by repeated squaring of z (a 6-deep dependent chain): each lane
// evaluates exp(-j*c*inc_t) directly from freq_offset (tau^c from a per-lane constant, the AFC part
// from an Estrin-form Taylor polynomial), so z and R_c are ready together.
#pragma once
#include "demod_core.cuh"
#include "fastmath.cuh"

namespace opvd {

constexpr int kWarpGateLaneE = 0, kWarpGateLaneO = 2, kWarpGateLaneL = 4;  // p of the gate owners
constexpr int kWarpLaneZ40 = 8;                                           // p whose R_c is z^40

// exp(-j*theta) for |theta| <= 0.4 from Taylor polynomials in Estrin form (truncation < 4e-17);
// sincos beyond (only reachable with a huge -o offset before the first AFC clamp, :303).
OPVD_HD cplx expmj(double th) {
    if (fabs(th) > 0.4) {
        double s, c;
        sincos(th, &s, &c);
        return {c, -s};
    }
    const double u = th * th;
    const double u2 = u * u;
    const double tu = th * u;
    // sin(th) = th + th*u*(s1 + s2 u + s3 u^2 + s4 u^3 + s5 u^4 + s6 u^5)
    const double sa = fma(u, 1.0 / 120.0, -1.0 / 6.0);
    const double sb = fma(u, 1.0 / 362880.0, -1.0 / 5040.0);
    const double sc = fma(u, 1.0 / 6227020800.0, -1.0 / 39916800.0);
    const double sp = fma(u2, fma(u2, sc, sb), sa);
    const double s = fma(tu, sp, th);
    // cos(th) = 1 + u*(c1 + c2 u + ... + c7 u^6)
    const double ca = fma(u, 1.0 / 24.0, -0.5);
    const double cb = fma(u, 1.0 / 40320.0, -1.0 / 720.0);
    const double cc = fma(u, 1.0 / 479001600.0, -1.0 / 3628800.0);
    const double cd = -1.0 / 87178291200.0;
    const double u4 = u2 * u2;
    const double cp = fma(u4, fma(u2, cd, cc), fma(u2, cb, ca));
    const double c = fma(u, cp, 1.0);
    return {c, -s};
}

// Per-lane constants and loop-carried registers.
struct WarpLane {
    int tone, p;
    double c;        // first slot of this lane (5p), as a double
    double sgn;      // -1 for F1 (inc = delta - 2pi/160), +1 for F2
    cplx tau1;       // exp(-j*sgn*2pi/160)       : z   = tau1 * exp(-j*delta)
    cplx tauc;       // exp(-j*sgn*2pi*c/160)     : R_c = tauc * exp(-j*c*delta)
    cplx z, R;       // current LO step and slot rotation (functions of freq_offset)
    double inc;      // LO phase step of this lane's tone (:210-211, :305-306)
    cplx prev;       // previous on-time correlation in the next symbol's phase frame (gate lane O only)
};

OPVD_HD void warp_lane_init(WarpLane& w, int lane, const FastMathTable& K) {
    w.tone = lane >> 4;
    w.p = lane & 15;
    const int pc = w.p > 12 ? 12 : w.p;
    w.c = 5.0 * pc;
    w.sgn = w.tone ? 1.0 : -1.0;
    w.tau1 = {K.tau_c, w.tone ? -K.tau_s : K.tau_s};
    // exp(-j*sgn*2*pi*c/160) with exact argument reduction: angle = pi * (5 pc) / 80
    double s, c;
#if defined(__CUDA_ARCH__)
    sincospi((5.0 * pc) / 80.0, &s, &c);
#else
    sincos(kPi * ((5.0 * pc) / 80.0), &s, &c);
#endif
    w.tauc = {c, w.tone ? -s : s};
    w.z = {1.0, 0.0};
    w.R = {1.0, 0.0};
    w.inc = 0.0;
    w.prev = {0.0, 0.0};
}

// LO step and slot rotation from the AFC offset (recomputed after every AFC update, :305-306)
OPVD_HD void warp_lane_lo(WarpLane& w, double freq_offset) {
    const double d = freq_offset * kTwoPiOverFs;
    w.inc = fma(w.sgn, kIncDev, d);
    w.z = cmul(w.tau1, expmj(d));
    w.R = cmul(w.tauc, expmj(w.c * d));
}

struct LanePartial {
    cplx W, F;
};

// I[0..4], Q[0..4]: samples of this lane's slots c..c+4
OPVD_HD LanePartial warp_lane_partial_d(const WarpLane& w, const double* I, const double* Q) {
    const cplx G = horner5(I, Q, w.z);
    LanePartial o;
    o.W = cmul(w.R, G);
    o.F = {fma(w.R.r, I[0], -(w.R.i * Q[0])), fma(w.R.r, Q[0], w.R.i * I[0])};
    return o;
}
// s[0..4]: the same samples packed
OPVD_HD LanePartial warp_lane_partial(const WarpLane& w, const uint32_t* s) {
    double I[5], Q[5];
#pragma unroll
    for (int r = 0; r < 5; ++r) unpack_iq(s[r], I[r], Q[r]);
    return warp_lane_partial_d(w, I, Q);
}

// LO step and slot rotation in the hot loop: valid for |freq_offset| <= 2.2 kHz, which the AFC clamp
// (:303) guarantees after every update.  Short Taylor polynomial for z, Estrin form for R_c.
OPVD_HD void warp_lane_lo_fast(WarpLane& w, double freq_offset, const FastMathTable& K) {
    const double d = freq_offset * K.two_pi_over_fs;
    w.inc = fma(w.sgn, K.inc_dev, d);
    w.z = cmul(w.tau1, expmj_small(d, K));
    w.R = cmul(w.tauc, expmj_mid(w.c * d, K));
}

// Interpolated gate on a gate-owner lane: C = sum of the 8 lane partials starting here,
// Fh = F of lane p+8, F = own F.  With dX = Fh - F (shifted-window edge term):
//   X = (1-f)*C + f*conj(z)*(C + dX) = C + f*(conj(z)*(C + dX) - C)
OPVD_HD cplx warp_lane_gate(const WarpLane& w, double f, cplx C, cplx Fh, cplx F) {
    const cplx S = {C.r + (Fh.r - F.r), C.i + (Fh.i - F.i)};
    const cplx T = {fma(w.z.r, S.r, w.z.i * S.i), fma(w.z.r, S.i, -(w.z.i * S.r))};  // conj(z) * S
    return {fma(f, T.r - C.r, C.r), fma(f, T.i - C.i, C.i)};
}

// AFC phase detector on gate lane O (:289-299).  X = interpolated O' (carrying z^10 = R of this lane),
// RP = R * prev, so X * conj(RP) = O_n * conj(prev).  corner: O_n or prev exactly zero (all-zero
// input) -> reproduce the reference's signed zeros through the slow path.
OPVD_HD_COLD double afc_phase_corner(cplx dom, cplx prev, double ph) { return afc_phase_signed_zero(dom, prev, ph); }

OPVD_HD double warp_lane_afc_phase(const WarpLane& w, cplx X, cplx RP, bool corner, double ph,
                                   const FastMathTable& K) {
    const double xr = fma(X.r, RP.r, X.i * RP.i);
    const double xi = fma(X.i, RP.r, -(X.r * RP.i));
    double pd = atan2_fast(xi, xr, K);  // branch-free; NaN in the corner case, replaced below
    // the rare fix-up comes AFTER the fast evaluation so that the common path stays one basic block
    // (the scheduler interleaves this chain with the timing loop's division)
    if (corner) pd = afc_phase_corner(cmul(X, cconj(w.R)), w.prev, ph);
    return pd;
}

// clamp to [-lim, lim] (same result as the reference's two comparisons for every non-NaN v)
OPVD_HD double clamp_sym(double v, double lim) { return fabs(v) > lim ? copysign(lim, v) : v; }

// early-gate correction for the first symbol of a call (:237), window read through an accessor
// (win(k) = packed raw sample of slot k); same value as first_symbol_fix() in demod_core.cuh.
template <class Win>
OPVD_HD cplx first_symbol_fix_w(Win win, double f, cplx z) {
    double I0, Q0;
    unpack_iq(win(kWinLead), I0, Q0);
    cplx fix = {0.0, 0.0};
    for (int k = 9; k >= 0; --k) {
        double Ia, Qa, Ib, Qb;
        unpack_iq(win(k), Ia, Qa);
        unpack_iq(win(k + 1), Ib, Qb);
        const double yr = fma(f, Ib - Ia, Ia) - I0;
        const double yi = fma(f, Qb - Qa, Qa) - Q0;
        fix = {fma(fix.r, z.r, fma(-fix.i, z.i, yr)), fma(fix.r, z.i, fma(fix.i, z.r, yi))};
    }
    return fix;
}

// AFC loop (:300-303) and LO phase wrap (:259-262) with table constants
OPVD_HD void warp_afc_loop(double& freq_offset, double pd, double afc_alpha, const FastMathTable& K) {
    const double ferr = pd * K.sym_rate_over_two_pi;
    freq_offset = clamp_sym(freq_offset + afc_alpha * ferr, 2000.0);
}
OPVD_HD double warp_wrap_phase(double ph, const FastMathTable& K) { return fma(-K.two_pi, rint(ph * K.inv_two_pi), ph); }

// Uniform (all lanes redundantly) soft decision, TED and timing loop (:264-286, :313).
// e1/e2 = |O|^2 per tone, eE*/eL* = early/late energies per tone.
OPVD_HD double warp_uniform_timing(double e1, double e2, double eE1, double eL1, double eE2, double eL2,
                                   double& timing_freq, double& pos, bool& tone1, const FastMathTable& K) {
    tone1 = e1 > e2;  // :272, :291
    const double ee = tone1 ? eE1 : eE2, el = tone1 ? eL1 : eL2;
    const double ted = div_fast(el - ee, el + ee + K.eps_ted);           // :280
    timing_freq = clamp_sym(timing_freq + K.k_tf * ted, K.lim_tf);       // :283-284
    const double adj = clamp_sym(K.k_adj * ted + timing_freq, 2.0);      // :285-286
    pos += 40.0 + adj;                                          // :313
    return e2 - e1;  // :268
}

}  // namespace opvd
