// oracle/ref_harness.cpp — TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Introspection harness around the UNMODIFIED reference receive chain.  The
// reference translation unit is pulled in *at build time* from where it lies
// (/root/reference/src/opv-demod.cpp, path given by -DOPV_REF_DEMOD_CPP=...) with
// its main() renamed; no reference source is copied into this repository.
// Output: oracle/_ref/libref_stages.so (git-ignored, travels to the GPU box).
//
// What it exposes (extern "C", plain pointers) are thin adapters that drive the
// reference's own classes so that tests can pin oracle/opv_oracle.c stage by stage:
//   MSKDemodulatorAFC::estimate_offset / demodulate   (src/opv-demod.cpp:131-329)
//   SyncTracker::process                               (src/opv-demod.cpp:615-736)
//   FrameDecoder::decode / ViterbiDecoder::decode      (src/opv-demod.cpp:802-898)
//   deinterleave_addr                                  (src/opv-demod.cpp:792-795)
// plus a restatement of the main() batch / streaming drivers (src/opv-demod.cpp:995-1216)
// that records soft symbols (the binary itself never prints them).

#include <vector>
#include <complex>
#include <array>
#include <string>
#include <cstdint>
#include <cstdio>
#include <cstring>

#define main opv_ref_main_unused
#include OPV_REF_DEMOD_CPP
#undef main

namespace {
std::vector<sample_t> to_cplx(const int16_t* iq, size_t n) {
    std::vector<sample_t> v(n);
    for (size_t i = 0; i < n; ++i) v[i] = sample_t(iq[2 * i], iq[2 * i + 1]);
    return v;
}
}  // namespace

extern "C" {

double ref_estimate_offset(const int16_t* iq, size_t n) {
    auto v = to_cplx(iq, n);
    MSKDemodulatorAFC d;
    return d.estimate_offset(v.data(), v.size());
}

// One demodulate() call on a fresh demodulator (batch semantics, optional preset offset).
size_t ref_demodulate_once(const int16_t* iq, size_t n, double freq_offset, double afc_alpha,
                           double* soft_out, size_t cap, double* final_freq, double* final_tfreq,
                           size_t* leftover) {
    auto v = to_cplx(iq, n);
    MSKDemodulatorAFC d;
    d.set_freq_offset(freq_offset);
    d.set_afc_bandwidth(afc_alpha);
    std::vector<double> soft;
    d.demodulate(v.data(), v.size(), soft);
    size_t m = soft.size() < cap ? soft.size() : cap;
    if (soft_out) memcpy(soft_out, soft.data(), m * sizeof(double));
    if (final_freq) *final_freq = d.get_freq_offset();
    if (final_tfreq) *final_tfreq = d.get_timing_freq();
    if (leftover) *leftover = d.get_leftover();
    return soft.size();
}

// Whole-capture soft symbols following main()'s two drivers.
//   streaming == 0 : src/opv-demod.cpp:1164-1174  (estimate on the whole capture, one demodulate call)
//   streaming == 1 : src/opv-demod.cpp:1012-1113  (86,720-sample chunks, leftover carry, EOF flush)
// chunk_starts (optional) receives the global sample index at which every demodulate() call began.
size_t ref_run_soft(const int16_t* iq, size_t n, int streaming, double afc_alpha,
                    int have_init_offset, double init_offset,
                    double* soft_out, size_t cap, double* est_offset_out, double* final_freq,
                    double* final_tfreq, int64_t* chunk_starts, size_t chunk_cap, size_t* n_chunks_out) {
    std::vector<double> all;
    MSKDemodulatorAFC demod;
    double est = 0.0;
    size_t n_chunks = 0;
    if (!streaming) {
        auto v = to_cplx(iq, n);
        est = demod.estimate_offset(v.data(), v.size());
        demod.set_freq_offset(est);
        demod.set_afc_bandwidth(afc_alpha);
        demod.demodulate(v.data(), v.size(), all);
        if (chunk_starts && chunk_cap) chunk_starts[0] = 0;
        n_chunks = 1;
    } else {
        if (have_init_offset) demod.set_freq_offset(init_offset);
        demod.set_afc_bandwidth(afc_alpha);
        const size_t CHUNK = FRAME_SYMBOLS * SAMPLES_PER_SYMBOL;
        std::vector<sample_t> buf;
        buf.reserve(CHUNK);
        bool first = true;
        size_t consumed_global = 0;  // global index of buf[0]
        for (size_t k = 0; k < n; ++k) {
            buf.push_back(sample_t(iq[2 * k], iq[2 * k + 1]));
            if (buf.size() >= CHUNK) {
                if (first) {
                    if (!have_init_offset) {
                        est = demod.estimate_offset(buf.data(), buf.size());
                        demod.set_freq_offset(est);
                    }
                    first = false;
                }
                if (chunk_starts && n_chunks < chunk_cap) chunk_starts[n_chunks] = (int64_t)consumed_global;
                ++n_chunks;
                std::vector<double> soft;
                demod.demodulate(buf.data(), buf.size(), soft);
                all.insert(all.end(), soft.begin(), soft.end());
                size_t leftover = demod.get_leftover();
                if (leftover > 0 && leftover < buf.size()) {
                    consumed_global += buf.size() - leftover;
                    std::vector<sample_t> keep(buf.end() - leftover, buf.end());
                    buf = std::move(keep);
                } else {
                    consumed_global += buf.size();
                    buf.clear();
                }
            }
        }
        if (!buf.empty()) {
            if (chunk_starts && n_chunks < chunk_cap) chunk_starts[n_chunks] = (int64_t)consumed_global;
            ++n_chunks;
            std::vector<double> soft;
            demod.demodulate(buf.data(), buf.size(), soft);
            all.insert(all.end(), soft.begin(), soft.end());
        }
    }
    size_t m = all.size() < cap ? all.size() : cap;
    if (soft_out) memcpy(soft_out, all.data(), m * sizeof(double));
    if (est_offset_out) *est_offset_out = est;
    if (final_freq) *final_freq = demod.get_freq_offset();
    if (final_tfreq) *final_tfreq = demod.get_timing_freq();
    if (n_chunks_out) *n_chunks_out = n_chunks;
    return all.size();
}

// Coherent mode (-c, batch only): src/opv-demod.cpp:1144-1161 on the reference's CoherentMSKDemodulator.
size_t ref_run_soft_coherent(const int16_t* iq, size_t n, double afc_alpha, double pll_bw, double* soft_out,
                             size_t cap, double* est_offset_out, double* final_freq) {
    auto v = to_cplx(iq, n);
    CoherentMSKDemodulator demod;
    double est = demod.estimate_offset(v.data(), v.size());
    demod.set_freq_offset(est);
    demod.set_afc_bandwidth(afc_alpha);
    demod.set_pll_bandwidth(pll_bw);
    std::vector<double> soft;
    demod.demodulate(v.data(), v.size(), soft);
    size_t m = soft.size() < cap ? soft.size() : cap;
    if (soft_out) memcpy(soft_out, soft.data(), m * sizeof(double));
    if (est_offset_out) *est_offset_out = est;
    if (final_freq) *final_freq = demod.get_freq_offset();
    return soft.size();
}

// SyncTracker + FrameDecoder over a soft-symbol sequence (src/opv-demod.cpp:1186-1205).
// The tracker's transition log goes to stderr exactly as in the reference binary.
// frame_ready_idx[k] = symbol index at which frame k became ready; metrics[k] = Viterbi metric
// (-1 frames are recorded with metric -1 and zero bytes so that drops are visible).
size_t ref_track_decode(const double* soft, size_t n, size_t idx0, uint8_t* frames, int32_t* metrics,
                        int64_t* frame_ready_idx, size_t cap, int* final_state) {
    SyncTracker tracker;
    FrameDecoder fdec;
    size_t nf = 0;
    for (size_t i = 0; i < n; ++i) {
        auto res = tracker.process(soft[i], idx0 + i);
        if (res.frame_ready && !res.payload.empty()) {
            std::array<uint8_t, FRAME_BYTES> frame{};
            int metric = fdec.decode(res.payload.data(), frame);
            if (nf < cap) {
                if (frames) memcpy(frames + nf * FRAME_BYTES, frame.data(), FRAME_BYTES);
                if (metrics) metrics[nf] = metric;
                if (frame_ready_idx) frame_ready_idx[nf] = (int64_t)(idx0 + i);
            }
            ++nf;
        }
    }
    if (final_state) *final_state = (int)tracker.get_state();
    return nf;
}

int ref_frame_decode(const double* soft2144, uint8_t* out134) {
    FrameDecoder fdec;
    std::array<uint8_t, FRAME_BYTES> frame{};
    int metric = fdec.decode(soft2144, frame);
    memcpy(out134, frame.data(), FRAME_BYTES);
    return metric;
}

int ref_viterbi_decode(const int32_t* soft_in2144, uint8_t* bits1072) {
    ViterbiDecoder v;
    std::array<int, ENCODED_BITS> in;
    for (size_t i = 0; i < ENCODED_BITS; ++i) in[i] = soft_in2144[i];
    std::array<uint8_t, FRAME_BITS> bits;
    int metric = v.decode(in, bits);
    memcpy(bits1072, bits.data(), FRAME_BITS);
    return metric;
}

size_t ref_deinterleave_addr(size_t idx) { return deinterleave_addr(idx); }

}  // extern "C"
