// fastmath.cuh — branch-free FP64 primitives for the per-symbol loop arithmetic of the demodulator
// (division for the timing error detector, atan2 for the AFC phase detector, exp(-j*theta) for the
// LO steps; reference call sites /root/reference/src/opv-demod.cpp:280, :299, :240-248).
//
// Why not the CUDA math library: in the latency-bound per-stream recurrence every instruction of the
// serial chain costs ~4 cycles.  libdevice's atan2/division carry slow-path branches and ~40 FP64
// immediates, each materialised with two UMOVs per use (ncu, profiles/: 150 of 555 warp-instructions
// per symbol were constant moves).  Here all coefficients live in one __constant__ table, so FP64
// instructions read them as constant-bank / uniform-register operands, polynomials are in Estrin
// form (depth ~5 instead of ~20), and there are no branches.  Accuracy: <= ~1 ulp-level (division)
// and < 2e-16 absolute (atan2, exp), i.e. the same class as libm-vs-libdevice differences that the
// original formulation already has against the reference.
// Host/device code (the host build is used by the CPU host-sim test only).
#pragma once
#include <cstring>

#include "opvd_common.cuh"

namespace opvd {

struct FastMathTable {
    double atan_c[12];    // P(u) ~ (atan(t)/t - 1)/u on u = t^2 in [0, tan(pi/8)^2], |err| < 2e-18
    double sin_c[6];      // -1/3!, 1/5!, ... (Taylor)
    double cos_c[7];      // -1/2!, 1/4!, ...
    double tan_pi_8;
    // loop constants of the demodulator that do not fit a 32-bit immediate
    double two_pi_over_fs, inc_dev, sym_rate_over_two_pi, two_pi, inv_two_pi;
    double eps_ted, k_tf, k_adj, lim_tf;  // 1e-10 (:280), 0.00001 (:283), 0.005 (:285), 0.1 (:284)
    double tau_c, tau_s;                  // cos, sin of 2*pi/160 (tone step)
    double pi_4;
};

#define OPVD_FASTMATH_TABLE_INIT                                                                          \
    {                                                                                                     \
        {-0.3333333333333333, 0.19999999999999762, -0.1428571428564902, 0.11111111104112661,              \
         -0.09090908702135936, 0.07692294904896305, -0.06666398746388635, 0.05878631460320977,            \
         -0.05228284767565546, 0.045417954500753324, -0.034321082112901316, 0.016012008547392122},        \
        {-1.0 / 6.0, 1.0 / 120.0, -1.0 / 5040.0, 1.0 / 362880.0, -1.0 / 39916800.0, 1.0 / 6227020800.0},  \
        {-0.5, 1.0 / 24.0, -1.0 / 720.0, 1.0 / 40320.0, -1.0 / 3628800.0, 1.0 / 479001600.0,              \
         -1.0 / 87178291200.0},                                                                           \
        0.41421356237309504880,                                                                           \
        2.8981482044186284e-06, 0.039269908169872414, 8626.197915580728, 6.283185307179586,               \
        0.15915494309189535, 1e-10, 0.00001, 0.005, 0.1, 0.9992290362407229, 0.03925981575906861,       \
        0.78539816339744830962                                                                            \
    }

#if defined(__CUDACC__)
static __constant__ FastMathTable g_fm = OPVD_FASTMATH_TABLE_INIT;
#else
static const FastMathTable g_fm = OPVD_FASTMATH_TABLE_INIT;
#endif

// Register-resident copy of the table for a long-running loop.  ptxas re-materialises values it can
// trace to an immediate or a constant-bank address inside the loop (2 UMOVs or an LDC per use); a
// volatile load from global memory cannot be repeated or traced, so the values stay in registers.
#if defined(__CUDACC__)
static __device__ FastMathTable g_fm_global = OPVD_FASTMATH_TABLE_INIT;
__device__ __forceinline__ FastMathTable load_table_pinned() {
    FastMathTable k;
    double* p = reinterpret_cast<double*>(&k);
    const volatile double* src = reinterpret_cast<const volatile double*>(&g_fm_global);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(FastMathTable) / sizeof(double)); ++i) p[i] = src[i];
    return k;
}
#endif

// seed of 1/b: MUFU.RCP64H on the device; the host stand-in keeps the upper mantissa bits of 1/b
OPVD_HD double rcp_seed(double b) {
#if defined(__CUDA_ARCH__)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
    return r;
#else
    double r = 1.0 / b;
    uint64_t u;
    std::memcpy(&u, &r, 8);
    u &= 0xFFFFFFFF00000000ull;  // keep ~20 mantissa bits, like the hardware seed
    std::memcpy(&r, &u, 8);
    return r;
#endif
}

// a / b for finite, normal b (no zero, subnormal, inf or NaN handling).  MUFU.RCP64H is accurate to
// 2^-20 (measured, profiles/microbench: max |1 - b*r| = 9.95e-7); one Newton step brings the
// reciprocal to 2^-40, and the residual correction of the quotient squares what is left, so the
// result is within ~1 ulp.  6 instructions, no branches (libdevice: 11 + a slow-path branch).
OPVD_HD double div_fast(double a, double b) {
    double r = rcp_seed(b);
    r = fma(r, fma(-b, r, 1.0), r);
    const double q = a * r;
    return fma(fma(-b, q, a), r, q);
}

OPVD_HD double flip_sign_if(double v, bool neg) {
#if defined(__CUDA_ARCH__)
    return __hiloint2double(__double2hiint(v) ^ (neg ? 0x80000000 : 0), __double2loint(v));
#else
    return neg ? -v : v;
#endif
}

// atan2(y, x) for finite inputs that are not both zero.
OPVD_HD double atan2_fast(double y, double x, const FastMathTable& K) {
    const double ax = fabs(x), ay = fabs(y);
    const bool swap = ay > ax;
    const double mx = swap ? ay : ax, mn = swap ? ax : ay;
    const bool big = mn > K.tan_pi_8 * mx;  // reduce to |t| <= tan(pi/8): atan(t) = pi/4 - atan((1-t)/(1+t))
    const double num = big ? mx - mn : mn, den = big ? mx + mn : mx;
    const double t = div_fast(num, den);
    const double u = t * t;
    const double u2 = u * u;
    const double u4 = u2 * u2;
    const double* c = K.atan_c;
    const double p01 = fma(u, c[1], c[0]), p23 = fma(u, c[3], c[2]), p45 = fma(u, c[5], c[4]);
    const double p67 = fma(u, c[7], c[6]), p89 = fma(u, c[9], c[8]), pab = fma(u, c[11], c[10]);
    const double q0 = fma(u2, p23, p01), q1 = fma(u2, p67, p45), q2 = fma(u2, pab, p89);
    const double P = fma(u4, fma(u4, q2, q1), q0);
    const double a0 = fma(t * u, P, t);
    // undo the reductions: a = m*pi/4 +/- a0 with m = big, then 2-m if swapped, then 4-m if x < 0
    // (integer selects, off the critical path; an indexed constant load here stalled the warp ~8 %)
    const bool xneg = x < 0.0;
    int m = big ? 1 : 0;
    m = swap ? 2 - m : m;
    m = xneg ? 4 - m : m;
    const double a = fma((double)m, K.pi_4, flip_sign_if(a0, (big != swap) != xneg));
    return copysign(a, y);
}

// exp(-j*theta), |theta| <= 0.0075 (LO step: |offset| <= 2.5 kHz): truncation < 1e-19
OPVD_HD cplx expmj_small(double th, const FastMathTable& K) {
    const double u = th * th;
    const double s = fma(th * u, fma(u, K.sin_c[1], K.sin_c[0]), th);
    const double c = fma(u, fma(u, fma(u, K.cos_c[2], K.cos_c[1]), K.cos_c[0]), 1.0);
    return {c, -s};
}

// exp(-j*theta), |theta| <= 0.4: Taylor polynomials in Estrin form, truncation < 4e-17
OPVD_HD cplx expmj_mid(double th, const FastMathTable& K) {
    const double u = th * th;
    const double u2 = u * u;
    const double* sc = K.sin_c;
    const double* cc = K.cos_c;
    const double sa = fma(u, sc[1], sc[0]), sb = fma(u, sc[3], sc[2]), sd = fma(u, sc[5], sc[4]);
    const double s = fma(th * u, fma(u2, fma(u2, sd, sb), sa), th);
    const double ca = fma(u, cc[1], cc[0]), cb = fma(u, cc[3], cc[2]), cd = fma(u, cc[5], cc[4]);
    const double u4 = u2 * u2;
    const double c = fma(u, fma(u4, fma(u2, cc[6], cd), fma(u2, cb, ca)), 1.0);
    return {c, -s};
}

}  // namespace opvd
