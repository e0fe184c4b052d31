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

}  // namespace opvd
