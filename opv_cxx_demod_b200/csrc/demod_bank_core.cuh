// demod_bank_core.cuh — arithmetic of the CHANNEL-BANK demodulator (kernels_demod_bank.cu), the kernel for
// thousands of streams.  Same algorithm as demod_core.cuh (reference: MSKDemodulatorAFC::demodulate,
// /root/reference/src/opv-demod.cpp:206-329); what changes is how little of it is evaluated and by whom.
// Host/device code: the CUDA kernel and the CPU host-sim test (tests/hostsim) compile exactly these functions.
//
// One lane owns one stream.  Per symbol the work of a stream is split between two roles that run as two warps
// on the same 32 streams and overlap in time:
//   WINDOW role  on-time block sums of both tones (window slots 10..49) -> soft decision and dominant tone;
//                THEN the early/late block sums (slots 0..9, 50..59) of the DOMINANT tone only -> TED -> timing
//                loop -> next position.  The reference evaluates early and late gates of both tones and uses
//                one pair (:271-280); skipping the other pair removes a sixth of the Horner work.
//   AFC role     phase detector (atan2), AFC loop, LO steps and their powers for the next symbol (:289-310).
// The AFC role starts as soon as the on-time sums exist, i.e. it runs concurrently with the window role's
// early/late + timing half, and the next symbol's Horner needs nothing but z from it.
//
// FP64 operation count per symbol and stream (the FP64 pipe is what bounds this kernel):
//   on-time Horner 4 blocks x 2 tones x 36 = 288, early/late Horner 2 x 36 = 72, Q conversions 61,
//   gate combination + interpolation ~100, timing ~25, AFC + LO ~115  =>  ~660, i.e. ~16.5 per input sample
//   (the minimum for six full gates is 12 per sample in Horner work alone; round 1's batched kernel spent ~25).
//
// LO powers.  z_t = tau^(+/-1) * zeta with tau = exp(j*2*pi/160) constant and zeta = exp(-j*delta) from the AFC
// offset, so z_t^k = tau^(+/-k) * zeta^k: one squaring chain of zeta serves both tones, tau^10, tau^20 are
// constants, and tau^40 = j exactly (z_1^40 = j*zeta^40, z_2^40 = -j*zeta^40).
#pragma once
#include "demod_core.cuh"
#include "fastmath.cuh"

namespace opvd {

struct BankConsts {
    double t10c, t10s;  // cos, sin of pi/8   (tau^10)
    double t20;         // cos = sin of pi/4  (tau^20)
};
#define OPVD_BANK_CONSTS_INIT {0.92387953251128674, 0.38268343236508977, 0.70710678118654752}
#if defined(__CUDACC__)
static __constant__ BankConsts g_bk = OPVD_BANK_CONSTS_INIT;
#else
static const BankConsts g_bk = OPVD_BANK_CONSTS_INIT;
#endif

OPVD_HD double clamp_sym_b(double v, double lim) { return fabs(v) > lim ? copysign(lim, v) : v; }
OPVD_HD_COLD double batch_afc_corner(cplx dom, cplx prev, double ph) { return afc_phase_signed_zero(dom, prev, ph); }

// a = (c + js) * v,  b = (c - js) * v   (6 operations for both)
OPVD_HD void tone_pair(double c, double s, cplx v, cplx& a, cplx& b) {
    const double m1 = c * v.r, m3 = c * v.i;
    a = {fma(-s, v.i, m1), fma(s, v.r, m3)};
    b = {fma(s, v.i, m1), fma(-s, v.r, m3)};
}

struct BankLo {
    cplx z1, z2;        // exp(-j*inc_t): the per-sample LO steps (:210-211, :305-306)
    double inc1, inc2;
};
struct BankPow {
    cplx q1, q2;        // z_t^10
    cplx qq1, qq2;      // z_t^20
    cplx zeta40;        // exp(-j*40*delta); z_1^40 = j*zeta40, z_2^40 = -j*zeta40
};
OPVD_HD cplx bank_z40(cplx zeta40, int tone) {  // tone 0 -> F1
    return tone ? cplx{zeta40.i, -zeta40.r} : cplx{-zeta40.i, zeta40.r};
}

OPVD_HD void bank_lo_from_zeta(cplx zeta, double d, BankLo& lo, const FastMathTable& K) {
    lo.inc1 = d - K.inc_dev;
    lo.inc2 = d + K.inc_dev;
    tone_pair(K.tau_c, K.tau_s, zeta, lo.z1, lo.z2);
}
OPVD_HD void bank_pow_from_zeta(cplx zeta, BankPow& p, const BankConsts& B) {
    const cplx a2 = csqr(zeta), a4 = csqr(a2), a5 = cmul(a4, zeta), a10 = csqr(a5), a20 = csqr(a10);
    p.zeta40 = csqr(a20);
    tone_pair(B.t10c, B.t10s, a10, p.q1, p.q2);
    tone_pair(B.t20, B.t20, a20, p.qq1, p.qq2);
}
// general offset (kernel start: a -o value may exceed the AFC clamp) / hot loop (|offset| <= 2 kHz after :303)
OPVD_HD cplx bank_zeta_general(double freq_offset, double& d) { return zeta_from_offset(freq_offset, d); }
OPVD_HD cplx bank_zeta_fast(double freq_offset, double& d, const FastMathTable& K) {
    d = freq_offset * K.two_pi_over_fs;
    return expmj_small(d, K);
}

// ---------------------------------------------------------------------------------------------------------
// WINDOW role.  win(k, I, Q) yields window slot k (raw sample b-10+k) as doubles.

// Horner step g = g*z + s
OPVD_HD void hstep(cplx& g, cplx z, double I, double Q) {
    const double nr = fma(g.r, z.r, fma(-g.i, z.i, I));
    const double ni = fma(g.r, z.i, fma(g.i, z.r, Q));
    g.r = nr; g.i = ni;
}

struct BankOnTime {
    cplx P1, R1, P2, R2;     // P = H1 + q*H2, R = H3 + q*H4 per tone (H_m = block sum of slots 10m..10m+9)
    cplx H2a, H3a, H2b, H3b; // block sums the early / late gates of the dominant tone still need (a: F1, b: F2)
    cplx s20, s40;           // raw samples of slots 20 and 40 (edge terms of the late / early gate)
    cplx O1, O2;             // interpolated on-time correlations (common unit-modulus phase factor dropped)
    double eO1, eO2;
};

// on-time block sums of both tones: 16 independent dependency chains (4 blocks x 2 tones x re/im).
// Register-file note (tools/microbench_rf.cu, tools/sass_stalls.py): on B200 a DFMA whose three 64-bit operands all come
// from the register file issues every 3.0 cycles, one with an operand held by the operand-reuse cache every 2.2.  The
// four DFMAs of a Horner step share z.r, z.i, g.r, g.i pairwise, so a schedule with one reused operand per DFMA exists,
// but ptxas interleaves the 16 chains its own way whatever the source order (level-by-level and rolled-loop forms were
// tried): 57 % of this kernel's FP64 instructions fetch three operands, which costs ~300 cycles per symbol.
template <class Win>
OPVD_HD void bank_on_blocks(Win win, cplx z1, cplx z2, cplx (&A)[4], cplx (&B)[4], cplx& s10, cplx& s20, cplx& s40) {
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        double I, Q;
        win(10 * (m + 1) + 9, I, Q);
        A[m] = {I, Q};
        B[m] = {I, Q};
    }
#pragma unroll
    for (int j = 8; j >= 0; --j) {
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            double I, Q;
            win(10 * (m + 1) + j, I, Q);
            hstep(A[m], z1, I, Q);
            hstep(B[m], z2, I, Q);
            if (j == 0 && m == 0) s10 = {I, Q};
            if (j == 0 && m == 1) s20 = {I, Q};
            if (j == 0 && m == 3) s40 = {I, Q};
        }
    }
}

// interpolated gate from the raw gate sum X and its shifted-window edge term dX = s[last+1]*z^40 - s[first]:
//   sum_k y[k] z^k = g*X + h*dX,   g = (1-f) + f*conj(z),  h = f*conj(z)     (demod_core.cuh)
OPVD_HD cplx bank_interp(cplx g, cplx h, cplx X, cplx last, cplx first, cplx z40) {
    const cplx dX = edge_term(last.r, last.i, first.r, first.i, z40);
    return cfma(g, X, cmul(h, dX));
}

template <class Win>
OPVD_HD void bank_on_time(Win win, double f, const BankLo& lo, const BankPow& pw, const cplx (&A)[4], const cplx (&B)[4],
                          cplx s10, cplx s20, cplx s40, BankOnTime& o) {
    o.P1 = cfma(pw.q1, A[1], A[0]); o.R1 = cfma(pw.q1, A[3], A[2]);
    o.P2 = cfma(pw.q2, B[1], B[0]); o.R2 = cfma(pw.q2, B[3], B[2]);
    const cplx X1 = cfma(pw.qq1, o.R1, o.P1), X2 = cfma(pw.qq2, o.R2, o.P2);
    o.H2a = A[1]; o.H3a = A[2]; o.H2b = B[1]; o.H3b = B[2];
    o.s20 = s20; o.s40 = s40;
    cplx s50;
    win(50, s50.r, s50.i);
    cplx g1, h1, g2, h2;
    interp_weights(lo.z1, f, g1, h1);
    interp_weights(lo.z2, f, g2, h2);
    o.O1 = bank_interp(g1, h1, X1, s50, s10, bank_z40(pw.zeta40, 0));
    o.O2 = bank_interp(g2, h2, X2, s50, s10, bank_z40(pw.zeta40, 1));
    o.eO1 = cnorm(o.O1);
    o.eO2 = cnorm(o.O2);
}

// early / late gates of the dominant tone (:271-280) -> their energies.  fixE: early-gate clamp of the first
// symbol of a call (:237), zero otherwise.
template <class Win>
OPVD_HD void bank_early_late(Win win, double f, bool tone1, const BankLo& lo, const BankPow& pw, const BankOnTime& o,
                             cplx fixE, double& eE, double& eL) {
    const cplx z = tone1 ? lo.z1 : lo.z2, q = tone1 ? pw.q1 : pw.q2, qq = tone1 ? pw.qq1 : pw.qq2;
    const cplx P = tone1 ? o.P1 : o.P2, R = tone1 ? o.R1 : o.R2;
    const cplx H2 = tone1 ? o.H2a : o.H2b, H3 = tone1 ? o.H3a : o.H3b;
    const cplx z40 = bank_z40(pw.zeta40, tone1 ? 0 : 1);
    // H0 = slots 0..9, H5 = slots 50..59
    cplx H0, H5, s0;
    {
        double I, Q;
        win(9, I, Q); H0 = {I, Q};
        win(59, I, Q); H5 = {I, Q};
    }
#pragma unroll
    for (int j = 8; j >= 0; --j) {
        double I, Q;
        win(j, I, Q);
        hstep(H0, z, I, Q);
        if (j == 0) s0 = {I, Q};
        win(50 + j, I, Q);
        hstep(H5, z, I, Q);
    }
    cplx s60;
    win(60, s60.r, s60.i);
    // E = H0 + q*H1 + q^2*H2 + q^3*H3 = H0 + q*(P + qq*H3);  L = H2 + q*H3 + q^2*H4 + q^3*H5 = H2 + q*(R + qq*H5)
    const cplx E = cfma(q, cfma(qq, H3, P), H0);
    const cplx L = cfma(q, cfma(qq, H5, R), H2);
    cplx g, h;
    interp_weights(z, f, g, h);
    cplx Ei = bank_interp(g, h, E, o.s40, s0, z40);
    const cplx Li = bank_interp(g, h, L, s60, o.s20, z40);
    Ei.r -= fixE.r; Ei.i -= fixE.i;
    eE = cnorm(Ei);
    eL = cnorm(Li);
}

// timing chain (:271-286, :313)
OPVD_HD void bank_timing(double eE, double eL, double& timing_freq, double& pos, const FastMathTable& K) {
    const double ted = div_fast(eL - eE, eL + eE + K.eps_ted);           // :280
    timing_freq = clamp_sym_b(timing_freq + K.k_tf * ted, K.lim_tf);     // :283-284
    const double adj = clamp_sym_b(K.k_adj * ted + timing_freq, 2.0);    // :285-286
    pos += 40.0 + adj;                                                   // :313
}

// ---------------------------------------------------------------------------------------------------------
// AFC role (:289-310, :250-262).  O1/O2: interpolated on-time sums of this symbol; zeta40 of the LO steps used
// DURING this symbol (previous correlations are stored rotated to the next symbol's phase frame).
struct BankAfc {
    double freq_offset, ph1, ph2;
    cplx p1, p2;
};
OPVD_HD void bank_afc(BankAfc& r, cplx O1, cplx O2, bool tone1, cplx zeta40, double inc1, double inc2, bool first_in_call,
                      double afc_alpha, const FastMathTable& K) {
    if (!first_in_call) {                                               // :289-307
        const cplx dom = tone1 ? O1 : O2, prev = tone1 ? r.p1 : r.p2;   // :291
        const double xr = fma(dom.r, prev.r, dom.i * prev.i);
        const double xi = fma(dom.i, prev.r, -(dom.r * prev.i));
        double pd = atan2_fast(xi, xr, K);
        const bool corner = (dom.r == 0.0 && dom.i == 0.0) || (prev.r == 0.0 && prev.i == 0.0);
        if (corner) pd = batch_afc_corner(dom, prev, tone1 ? r.ph1 : r.ph2);
        const double ferr = pd * K.sym_rate_over_two_pi;
        r.freq_offset = clamp_sym_b(r.freq_offset + afc_alpha * ferr, 2000.0);
    }
    // previous correlations for the NEXT symbol (:309-310): O * conj(z^40), z_1^40 = j*zeta40, z_2^40 = -j*zeta40
    r.p1 = cmul(O1, cconj(bank_z40(zeta40, 0)));
    r.p2 = cmul(O2, cconj(bank_z40(zeta40, 1)));
    const double a1 = fma(40.0, inc1, r.ph1), a2 = fma(40.0, inc2, r.ph2);  // :250-262
    r.ph1 = fma(-K.two_pi, rint(a1 * K.inv_two_pi), a1);
    r.ph2 = fma(-K.two_pi, rint(a2 * K.inv_two_pi), a2);
}

// ---------------------------------------------------------------------------------------------------------
// Early / late work split (ELB) for banks that leave SM sub-partitions idle (at most one 32-stream CTA per SM): the AFC
// warp evaluates the block sums H0 (slots 0..9) and H5 (slots 50..59) of BOTH tones while the window warp is still busy
// with the on-time blocks, and the window warp only combines the blocks of the dominant tone.  Same values as
// bank_early_late (a block sum does not depend on which warp evaluates it); 72 more DFMAs per symbol, ~190 fewer
// instructions on the critical warp: 10 % faster at one CTA per SM, 10 % slower at two and 5 % slower at four (the AFC
// warp then shares its sub-partition with another CTA's window warp).
struct BankElBlocks {
    cplx H0a, H0b, H5a, H5b;  // a: F1, b: F2
    cplx s0, s60;
};
template <class Win>
OPVD_HD void bank_el_blocks(Win win, cplx z1, cplx z2, BankElBlocks& e) {
    double I, Q;
    win(9, I, Q);  e.H0a = {I, Q}; e.H0b = {I, Q};
    win(59, I, Q); e.H5a = {I, Q}; e.H5b = {I, Q};
#pragma unroll
    for (int j = 8; j >= 0; --j) {
        win(j, I, Q);
        hstep(e.H0a, z1, I, Q);
        hstep(e.H0b, z2, I, Q);
        if (j == 0) e.s0 = {I, Q};
        win(50 + j, I, Q);
        hstep(e.H5a, z1, I, Q);
        hstep(e.H5b, z2, I, Q);
    }
    win(60, e.s60.r, e.s60.i);
}
OPVD_HD void bank_early_late_from_blocks(double f, bool tone1, const BankLo& lo, const BankPow& pw, const BankOnTime& o,
                                         cplx H0, cplx H5, cplx s0, cplx s60, cplx fixE, double& eE, double& eL) {
    const cplx z = tone1 ? lo.z1 : lo.z2, q = tone1 ? pw.q1 : pw.q2, qq = tone1 ? pw.qq1 : pw.qq2;
    const cplx P = tone1 ? o.P1 : o.P2, R = tone1 ? o.R1 : o.R2;
    const cplx H2 = tone1 ? o.H2a : o.H2b, H3 = tone1 ? o.H3a : o.H3b;
    const cplx z40 = bank_z40(pw.zeta40, tone1 ? 0 : 1);
    const cplx E = cfma(q, cfma(qq, H3, P), H0);
    const cplx L = cfma(q, cfma(qq, H5, R), H2);
    cplx g, h;
    interp_weights(z, f, g, h);
    cplx Ei = bank_interp(g, h, E, o.s40, s0, z40);
    const cplx Li = bank_interp(g, h, L, s60, o.s20, z40);
    Ei.r -= fixE.r; Ei.i -= fixE.i;
    eE = cnorm(Ei);
    eL = cnorm(Li);
}

}  // namespace opvd
