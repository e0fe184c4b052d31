// opvd_common.cuh — constants and small helpers shared by every kernel of the B200 opv-demod path.
//
// Signal constants follow the reference (/root/reference/src/opv-demod.cpp:39-60):
// 40 samples/symbol at 2.168 MSPS, tones at -/+13,550 Hz, 24-bit sync 0x02B8DB,
// 134-byte frames = 1072 bits -> 2144 coded bits, 2168 symbols per frame.
#pragma once
#include <cstdint>
#include <cmath>

#if defined(__CUDACC__)
#define OPVD_HD __host__ __device__ __forceinline__
#define OPVD_D __device__ __forceinline__
#define OPVD_HD_COLD static __host__ __device__ __noinline__
#else
#define OPVD_HD inline
#define OPVD_D inline
#define OPVD_HD_COLD static inline
#endif

namespace opvd {

constexpr int kSps = 40;                      // :39
constexpr double kSampleRate = 2168000.0;     // :40
constexpr double kSymbolRate = 54200.0;       // :41
constexpr double kFreqDev = 13550.0;          // :42
constexpr double kPi = 3.14159265358979323846;  // :43
constexpr double kTwoPi = 2.0 * kPi;          // :44
constexpr uint32_t kSyncWord = 0x02B8DB;      // :46
constexpr int kSyncBits = 24;                 // :47
constexpr int kFrameBytes = 134;              // :49
constexpr int kFrameBits = 1072;              // :50
constexpr int kEncodedBits = 2144;            // :51
constexpr int kFrameSymbols = 2168;           // :52
constexpr int kSyncMissLimit = 5;             // :60
constexpr int64_t kChunkSamples = 86720;      // :1012
constexpr int kEstSamples = 40000;            // :141 (SAMPLES_PER_SYMBOL * 1000)

// window of raw samples one symbol touches: local indices b-10 .. b+50 (early gate, on-time,
// late gate, +1 for the linear interpolator), b = floor(pos)
constexpr int kWin = 61;
constexpr int kWinLead = 10;

enum Mode : int32_t { kModeBatch = 0, kModeStream = 1 };
enum SyncState : int32_t { kHunting = 0, kVerifying = 1, kLocked = 2 };
enum EventType : int32_t { kEvHuntToVerify = 1, kEvVerifyToLocked = 2, kEvSyncOk = 3, kEvSyncMiss = 4, kEvLostLock = 5 };

struct cplx {
    double r, i;
};

OPVD_HD cplx cmul(cplx a, cplx b) { return {fma(a.r, b.r, -(a.i * b.i)), fma(a.r, b.i, a.i * b.r)}; }
OPVD_HD cplx csqr(cplx a) { return {fma(a.r, a.r, -(a.i * a.i)), (a.r + a.r) * a.i}; }
// a*b + c
OPVD_HD cplx cfma(cplx a, cplx b, cplx c) {
    return {fma(a.r, b.r, fma(-a.i, b.i, c.r)), fma(a.r, b.i, fma(a.i, b.r, c.i))};
}
OPVD_HD cplx cconj(cplx a) { return {a.r, -a.i}; }
OPVD_HD double cnorm(cplx a) { return fma(a.r, a.r, a.i * a.i); }

OPVD_HD double clampd(double v, double lo, double hi) { return (v < lo) ? lo : ((hi < v) ? hi : v); }

// packed sample word (I in bits 0..15, Q in bits 16..31, both signed) -> doubles (exact).
// On sm_100a each half converts with ONE instruction straight from the packed register
// (I2F.F64.S16 R, Rw / I2F.F64.S16 R, Rw.H1); OPVD_CVT_MAGIC selects the alternative
// 2^52-bias trick (LOP3 + DADD on the FP64 pipe) for comparison.
OPVD_HD void unpack_iq(uint32_t w, double& I, double& Q) {
#if defined(__CUDA_ARCH__) && defined(OPVD_CVT_MAGIC)
    I = __hiloint2double(0x43300000, (int)((w & 0xFFFFu) ^ 0x8000u)) - 4503599627403264.0;  // 2^52 + 2^15
    Q = __hiloint2double(0x43300000, (int)((w >> 16) ^ 0x8000u)) - 4503599627403264.0;
#else
    I = (double)(int16_t)(w & 0xFFFFu);
    Q = (double)(int16_t)(w >> 16);
#endif
}

// Same values, conversions split over two pipes: I through I2F.F64.S16 (XU pipe, 8 cycles per warp
// instruction), Q through the 2^52 bias trick (integer pipe + one DADD on the FP64 pipe).  A kernel
// that converts 60 components per thread and symbol is otherwise bound by the XU pipe.
OPVD_HD void unpack_iq_mixed(uint32_t w, double& I, double& Q) {
#if defined(__CUDA_ARCH__)
    I = (double)(int16_t)(w & 0xFFFFu);
    Q = __hiloint2double(0x43300000, (int)((w >> 16) ^ 0x8000u)) - 4503599627403264.0;  // 2^52 + 2^15
#else
    unpack_iq(w, I, Q);
#endif
}

}  // namespace opvd
