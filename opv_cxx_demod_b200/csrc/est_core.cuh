// est_core.cuh — coarse carrier-offset estimate (reference: MSKDemodulatorAFC::estimate_offset,
// /root/reference/src/opv-demod.cpp:131-202), restructured.
//
// The reference evaluates, for 121 coarse + 13 fine candidate offsets o,
//     E(o) = sum_{sym<1000} |sum_{i<40} s[40 sym+i] e^{-j phi1}|^2 + |... e^{-j phi2}|^2
// with running phases; only block energies are summed, so the phase at each block start is
// irrelevant and E(o) = P(w1(o)) + P(w2(o)) with P the summed 40-point periodogram
//     P(w) = R[0] + 2 sum_{l=1..39} ( Rr[l] cos(w l) + Ri[l] sin(w l) ),
//     R[l] = sum_blocks sum_{i=0}^{39-l} s[i+l] conj(s[i])      (block autocorrelation)
// R[l] is an exact integer (|R| < 2^47), accumulated exactly in FP64 FMAs from the int16 samples;
// every candidate then costs 39 sincos instead of 160,000 trig calls.  Candidate offsets are
// multiples of 5 Hz, so w*l is reduced exactly in integer arithmetic (units of 1/433600 turn).
// Candidate order, the strict '>' comparisons and the coarse->fine rule are the reference's.
#pragma once
#include "opvd_common.cuh"

namespace opvd {

constexpr int kEstLags = 40;

OPVD_HD void sincos_turns_433600(long long k, double* s, double* c) {
    // angle = 2*pi*k/433600, k reduced to [0, 433600)
    k %= 433600; if (k < 0) k += 433600;
#if defined(__CUDA_ARCH__)
    sincospi((double)(2 * k) / 433600.0, s, c);
#else
    sincos(kTwoPi * ((double)k / 433600.0), s, c);
#endif
}

// offset_hz must be a multiple of 5 Hz
OPVD_HD double est_energy(const double* Rr, const double* Ri, double offset_hz) {
    const long long m1 = (long long)((offset_hz - kFreqDev) / 5.0);  // exact: both multiples of 5
    const long long m2 = (long long)((offset_hz + kFreqDev) / 5.0);
    double acc = 0.0;
    for (int l = kEstLags - 1; l >= 1; --l) {
        double s1, c1, s2, c2;
        sincos_turns_433600(m1 * l, &s1, &c1);
        sincos_turns_433600(m2 * l, &s2, &c2);
        acc += fma(Rr[l], c1 + c2, Ri[l] * (s1 + s2));
    }
    return 2.0 * (acc + Rr[0]);
}

// sequential candidate scan exactly as :133-201 given an energy functor E(offset)
template <class EnergyFn>
OPVD_HD double est_scan(EnergyFn E) {
    double best_offset = 0, best_energy = 0;
    for (double offset = -1500; offset <= 1500; offset += 25) {
        double e = E(offset);
        if (e > best_energy) { best_energy = e; best_offset = offset; }
    }
    double fine_best = best_offset;
    for (double offset = best_offset - 30; offset <= best_offset + 30; offset += 5) {
        double e = E(offset);
        if (e > best_energy) { best_energy = e; fine_best = offset; }
    }
    return fine_best;
}

}  // namespace opvd
