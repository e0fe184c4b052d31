// Arithmetic of the bank-kernel variants that lost: on-time blocks in two passes (OB2), the window cut in
// two halves for two dependent window warps (LO/HI), early/late block sums on the AFC warp (ELB).  Each computes the same
// values as demod_bank_core.cuh.  Not part of the product build.
#pragma once
#include "demod_bank_core.cuh"
namespace opvd {

// The same sums in two passes of two blocks (8 chains instead of 16) with the samples of step j-1 fetched and converted
// before the Horner products of step j are written down: fewer live accumulators, deeper conversion prefetch.  Same
// values (a block sum does not depend on what is interleaved with it).
template <class Win>
OPVD_HD void bank_on_blocks_2x2(Win win, cplx z1, cplx z2, cplx (&A)[4], cplx (&B)[4], cplx& s10, cplx& s20, cplx& s40) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        double I0, Q0, I1, Q1;
        win(10 * (2 * h + 1) + 9, I0, Q0);
        win(10 * (2 * h + 2) + 9, I1, Q1);
        A[2 * h] = {I0, Q0}; B[2 * h] = {I0, Q0};
        A[2 * h + 1] = {I1, Q1}; B[2 * h + 1] = {I1, Q1};
        win(10 * (2 * h + 1) + 8, I0, Q0);
        win(10 * (2 * h + 2) + 8, I1, Q1);
#pragma unroll
        for (int j = 8; j >= 0; --j) {
            double nI0 = 0.0, nQ0 = 0.0, nI1 = 0.0, nQ1 = 0.0;
            if (j > 0) {
                win(10 * (2 * h + 1) + j - 1, nI0, nQ0);
                win(10 * (2 * h + 2) + j - 1, nI1, nQ1);
            }
            hstep(A[2 * h], z1, I0, Q0);
            hstep(B[2 * h], z2, I0, Q0);
            hstep(A[2 * h + 1], z1, I1, Q1);
            hstep(B[2 * h + 1], z2, I1, Q1);
            if (j == 0) {
                if (h == 0) { s10 = {I0, Q0}; s20 = {I1, Q1}; }
                else s40 = {I1, Q1};
            }
            I0 = nI0; Q0 = nQ0; I1 = nI1; Q1 = nQ1;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// The same window work cut in two halves for the four-warp kernel (two window warps per 32 streams):
//   LO warp: blocks H1, H2 (slots 10..29) -> P_t;  on-time combination, soft decision, dominant tone; H0 -> early gate
//   HI warp: blocks H3, H4 (slots 30..49) -> R_t;  H5 of both tones (no need to wait for the decision) -> late gate
// Every value is computed by exactly the operations of bank_on_blocks / bank_on_time / bank_early_late above, so
// the two kernels produce identical bits.

// two adjacent 10-slot blocks starting at slot k0, both tones; fA/fB = first raw sample of each block
template <class Win>
OPVD_HD void bank_two_blocks(Win win, int k0, cplx z1, cplx z2, cplx (&A)[2], cplx (&B)[2], cplx& fA, cplx& fB) {
#pragma unroll
    for (int m = 0; m < 2; ++m) {
        double I, Q;
        win(k0 + 10 * m + 9, I, Q);
        A[m] = {I, Q};
        B[m] = {I, Q};
    }
#pragma unroll
    for (int j = 8; j >= 0; --j) {
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            double I, Q;
            win(k0 + 10 * m + j, I, Q);
            hstep(A[m], z1, I, Q);
            hstep(B[m], z2, I, Q);
            if (j == 0 && m == 0) fA = {I, Q};
            if (j == 0 && m == 1) fB = {I, Q};
        }
    }
}
// one 10-slot block starting at slot k0, both tones
template <class Win>
OPVD_HD void bank_block_both(Win win, int k0, cplx z1, cplx z2, cplx& H1, cplx& H2) {
    double I, Q;
    win(k0 + 9, I, Q);
    H1 = {I, Q};
    H2 = {I, Q};
#pragma unroll
    for (int j = 8; j >= 0; --j) {
        win(k0 + j, I, Q);
        hstep(H1, z1, I, Q);
        hstep(H2, z2, I, Q);
    }
}
// one 10-slot block starting at slot k0, one tone; f0 = its first raw sample
template <class Win>
OPVD_HD void bank_block_one(Win win, int k0, cplx z, cplx& H, cplx& f0) {
    double I, Q;
    win(k0 + 9, I, Q);
    H = {I, Q};
#pragma unroll
    for (int j = 8; j >= 0; --j) {
        win(k0 + j, I, Q);
        hstep(H, z, I, Q);
        if (j == 0) f0 = {I, Q};
    }
}
// LO warp: on-time correlations from its own P_t and the HI warp's R_t
OPVD_HD void bank_on_time_from_halves(double f, const BankLo& lo, const BankPow& pw, cplx P1, cplx P2, cplx R1, cplx R2, cplx s10,
                                      cplx s50, cplx& O1, cplx& O2, double& eO1, double& eO2) {
    const cplx X1 = cfma(pw.qq1, R1, P1), X2 = cfma(pw.qq2, R2, P2);
    cplx g1, h1, g2, h2;
    interp_weights(lo.z1, f, g1, h1);
    interp_weights(lo.z2, f, g2, h2);
    O1 = bank_interp(g1, h1, X1, s50, s10, bank_z40(pw.zeta40, 0));
    O2 = bank_interp(g2, h2, X2, s50, s10, bank_z40(pw.zeta40, 1));
    eO1 = cnorm(O1);
    eO2 = cnorm(O2);
}
// interpolated gate energy of  G = a + q*(b + qq*c)  (early: a = H0, b = P, c = H3; late: a = H2, b = R, c = H5)
OPVD_HD double bank_gate_energy(double f, cplx z, cplx q, cplx qq, cplx z40, cplx a, cplx b, cplx c, cplx last, cplx first,
                                cplx fix) {
    const cplx G = cfma(q, cfma(qq, c, b), a);
    cplx g, h;
    interp_weights(z, f, g, h);
    cplx Gi = bank_interp(g, h, G, last, first, z40);
    Gi.r -= fix.r; Gi.i -= fix.i;
    return cnorm(Gi);
}

// ---------------------------------------------------------------------------------------------------------
// Early / late work split for the kernel variant in which the AFC warp evaluates the block sums H0 (slots 0..9) and
// H5 (slots 50..59) of BOTH tones while the window warp is still busy with the on-time blocks: the window warp then
// only combines the blocks of the dominant tone.  Same values as bank_early_late (a block sum does not depend on
// which warp evaluates it); 72 more DFMAs per symbol, ~190 fewer instructions on the critical warp.
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
