// demod_pipe_core.cuh — window arithmetic of the PIPELINED demodulator (kernels_demod_pipe.cu): the 61-slot
// window of one symbol is split into four QUARTERS of 15 slots, each handled by one thread for BOTH tones, so
// that every raw sample is loaded and converted once per symbol (the batched kernel's tone x half split
// converts it twice).  Same algorithm as demod_core.cuh / demod_batch_core.cuh (reference:
// MSKDemodulatorAFC::demodulate, /root/reference/src/opv-demod.cpp:206-329; the six gate correlations of
// :233-248 as polynomials in z = exp(-j*inc) over the raw samples, the linear interpolator of :122-128
// applied after the sums).  Host/device code: the CUDA kernel and the CPU host-sim test compile exactly these
// functions.
//
// Window slots k = 0..60 are the raw samples b-10 .. b+50.  5-slot blocks and q5 = z^5:
//   G_m = sum_{j<5} s[5m+j] z^j,  m = 0..11                                  (Horner, quarter threads)
//   E = sum_{m=0..7} q5^m G_m,   O = sum_{m=2..9} q5^(m-2) G_m,   L = sum_{m=4..11} q5^(m-4) G_m
// Quarter w owns blocks 3w..3w+2 and hands over two partial sums P_w, R_w:
//   w = 0:  P = G0 + q5 G1 + q5^2 G2  (E)          R = G2                       (O)
//   w = 1:  P = G3 + q5 G4 + q5^2 G5  (E and O)    R = G4 + q5 G5               (L)
//   w = 2:  P = G6 + q5 G7            (E)          R = G6 + q5 G7 + q5^2 G8     (O and L)
//   w = 3:  P = G9                    (O)          R = G9 + q5 G10 + q5^2 G11   (L)
//   E = P0 + q5^3 P1 + q5^6 P2,   O = R0 + q5 P1 + q5^4 R2 + q5^7 P3,   L = R1 + q5^2 R2 + q5^5 R3
// The post-sum interpolator X = C + f*(conj(z)*(C + dX) - C) with the edge term dX = s[last+1] z^40 - s[first]
// (dE: slots 40 and 0, dO: 50 and 10, dL: 60 and 20) is applied once per gate by the finishing thread.
#pragma once
#include "demod_batch_core.cuh"

namespace opvd {

// (horner5: demod_core.cuh)

struct QuarterParts {
    cplx P, R;
};

// quarter w of the window, one tone: I/Q = the quarter's 15 samples (slots 15w .. 15w+14)
OPVD_HD QuarterParts quarter_parts(const double* I, const double* Q, cplx z, cplx q5, int w) {
    const cplx A = horner5(I, Q, z), B = horner5(I + 5, Q + 5, z), C = horner5(I + 10, Q + 10, z);
    QuarterParts o;
    if (w == 0) {
        o.P = cfma(q5, cfma(q5, C, B), A);
        o.R = C;
    } else if (w == 1) {
        o.R = cfma(q5, C, B);
        o.P = cfma(q5, o.R, A);
    } else if (w == 2) {
        o.P = cfma(q5, B, A);
        o.R = cfma(q5, cfma(q5, C, B), A);
    } else {
        o.P = A;
        o.R = cfma(q5, cfma(q5, C, B), A);
    }
    return o;
}

// Finishing one tone: the four quarters' partial sums, the six raw edge samples (slots 0, 10, 20, 40, 50, 60 as
// complex doubles), the tone's LO steps (z, z5 = z^5, q = z^10) and the interpolation fraction.  Split into two
// independent halves so that two threads can share a tone: the early/late energies, and the on-time sum.
// fixE: early-gate clamp of the first symbol of a call (:237), zero otherwise.
OPVD_HD cplx quarter_interp(cplx Cg, cplx s_last, cplx s_first, cplx z, cplx z40, double f) {
    // S = Cg + s_last * z^40 - s_first;  X = Cg + f * (conj(z) * S - Cg)
    const cplx S = {fma(s_last.r, z40.r, fma(-s_last.i, z40.i, Cg.r - s_first.r)),
                    fma(s_last.r, z40.i, fma(s_last.i, z40.r, Cg.i - s_first.i))};
    const cplx T = {fma(z.r, S.r, z.i * S.i), fma(z.r, S.i, -(z.i * S.r))};
    return cplx{fma(f, T.r - Cg.r, Cg.r), fma(f, T.i - Cg.i, Cg.i)};
}
struct EarlyLate {
    double eE, eL;
};
OPVD_HD EarlyLate finish_early_late(const QuarterParts (&p)[4], const cplx (&edge)[6], const ToneLo& t, double f,
                                    cplx fixE) {
    const cplx q1 = t.z5, q2 = t.q;
    const cplx q3 = cmul(q2, q1), q4 = csqr(q2);
    const cplx q5 = cmul(q4, q1), q6 = csqr(q3), q8 = csqr(q4);  // q8 = z^40
    const cplx E = cfma(q6, p[2].P, cfma(q3, p[1].P, p[0].P));
    const cplx L = cfma(q5, p[3].R, cfma(q2, p[2].R, p[1].R));
    cplx Ei = quarter_interp(E, edge[3], edge[0], t.z, q8, f);
    const cplx Li = quarter_interp(L, edge[5], edge[2], t.z, q8, f);
    Ei.r -= fixE.r; Ei.i -= fixE.i;
    return {cnorm(Ei), cnorm(Li)};
}
struct OnTimeGate {
    cplx O, z40;
    double eO;
};
OPVD_HD OnTimeGate finish_on_time(const QuarterParts (&p)[4], const cplx (&edge)[6], const ToneLo& t, double f) {
    const cplx q1 = t.z5, q2 = t.q;
    const cplx q3 = cmul(q2, q1), q4 = csqr(q2);
    const cplx q7 = cmul(q4, q3), q8 = csqr(q4);
    const cplx O = cfma(q7, p[3].P, cfma(q4, p[2].R, cfma(q1, p[1].P, p[0].R)));
    OnTimeGate o;
    o.O = quarter_interp(O, edge[4], edge[1], t.z, q8, f);
    o.eO = cnorm(o.O);
    o.z40 = q8;
    return o;
}
OPVD_HD ToneGates finish_tone_quarters(const QuarterParts (&p)[4], const cplx (&edge)[6], const ToneLo& t, double f,
                                       cplx fixE) {
    const EarlyLate el = finish_early_late(p, edge, t, f, fixE);
    const OnTimeGate ot = finish_on_time(p, edge, t, f);
    ToneGates o;
    o.O = ot.O; o.eO = ot.eO; o.z40 = ot.z40; o.eE = el.eE; o.eL = el.eL;
    return o;
}

}  // namespace opvd
