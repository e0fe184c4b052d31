// opv_demod_main.cpp — drop-in `opv-demod` executable over the libopvd C ABI.
//
// Process contract of the reference (/root/reference/src/opv-demod.cpp:943-1217):
//   argv    -q quiet, -r raw frames on stdout, -s streaming, -a <alpha>, -o <hz> (streaming only),
//           -c coherent mode and -p <hz> PLL bandwidth (batch only, like the reference: with -s the streaming
//           branch returns first), -h help; unknown arguments ignored
//   stdin   int16 LE interleaved I/Q, read to EOF; a trailing partial sample is dropped
//   stdout  with -r: 134-byte frames, one write + flush per frame
//   stderr  banner, tracker transitions (always), frame boxes and summary unless -q
//   exit    0 iff at least one frame was decoded
// All signal processing runs on the GPU through libopvd; this file only moves bytes and prints.
#include <unistd.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/opvd.h"

namespace {

const char* state_name(int s) {
    switch (s) {
        case OPVD_HUNTING: return "HUNTING";
        case OPVD_VERIFYING: return "VERIFYING";
        case OPVD_LOCKED: return "LOCKED";
    }
    return "?";
}

// Base-40 station id, first character in the least significant digit (:87-103)
std::string base40(const uint8_t* b) {
    uint64_t v = 0;
    for (int i = 0; i < 6; ++i) v = (v << 8) | b[i];
    if (!v) return "(empty)";
    std::string out;
    for (; v > 0; v /= 40) {
        const int d = (int)(v % 40);
        char c = 0;
        if (d >= 1 && d <= 26) c = (char)('A' + d - 1);
        else if (d >= 27 && d <= 36) c = (char)('0' + d - 27);
        else if (d == 37) c = '-';
        else if (d == 38) c = '/';
        else if (d == 39) c = '.';
        if (c) out += c;
    }
    return out.empty() ? "(empty)" : out;
}

void print_frame(int num, const uint8_t* f, int metric, double sync) {  // layout of :907-938
    const char* bar = "─────────────────────────────────────────────────────────────────";
    fprintf(stderr, "┌%s┐\n", bar);
    fprintf(stderr, "│ FRAME %4d  │  Sync: %.3f  │  Metric: %5d", num, sync, metric);
    if (metric == 0) fprintf(stderr, " (perfect)");
    fprintf(stderr, "\n├%s┤\n", bar);
    fprintf(stderr, "│ Station ID:  %-12s (Base-40)\n", base40(f).c_str());
    const uint32_t tok = (f[6] << 16) | (f[7] << 8) | f[8];
    fprintf(stderr, "│ Token:       0x%06X%s\n", tok, tok == 0xBBAADD ? " (default)" : "");
    fprintf(stderr, "│ Reserved:    0x%06X\n", (unsigned)((f[9] << 16) | (f[10] << 8) | f[11]));
    fprintf(stderr, "├%s┤\n", bar);
    fprintf(stderr, "│ Hex Dump:                                                       │\n");
    for (int i = 0; i < OPVD_FRAME_BYTES; i += 16) {
        fprintf(stderr, "│ %02zx: ", (size_t)i);
        for (int j = i; j < i + 16 && j < OPVD_FRAME_BYTES; ++j) fprintf(stderr, "%02X ", f[j]);
        for (int j = OPVD_FRAME_BYTES; j < i + 16; ++j) fprintf(stderr, "   ");
        fprintf(stderr, " │");
        for (int j = i; j < i + 16 && j < OPVD_FRAME_BYTES; ++j) fprintf(stderr, "%c", (f[j] >= 0x20 && f[j] < 0x7F) ? f[j] : '.');
        fprintf(stderr, "│\n");
    }
    fprintf(stderr, "└%s┘\n\n", bar);
}

void print_event(const opvd_event& e) {  // :651,:677,:695,:699,:705
    switch (e.type) {
        case OPVD_EV_HUNT_TO_VERIFY:
            fprintf(stderr, "[%zu] HUNTING→VERIFYING (corr=%.3f, raw=%.0f)\n", (size_t)e.sym_idx, e.corr, e.raw);
            break;
        case OPVD_EV_VERIFY_TO_LOCKED:
            fprintf(stderr, "[%zu] VERIFYING→LOCKED (frame %d)\n", (size_t)e.sym_idx, e.count);
            break;
        case OPVD_EV_SYNC_OK: fprintf(stderr, "[%zu] LOCKED: sync OK (corr=%.3f)\n", (size_t)e.sym_idx, e.corr); break;
        case OPVD_EV_SYNC_MISS:
            fprintf(stderr, "[%zu] LOCKED: sync MISS #%d (corr=%.3f)\n", (size_t)e.sym_idx, e.count, e.corr);
            break;
        case OPVD_EV_LOST_LOCK: fprintf(stderr, "[%zu] LOCKED→HUNTING (lost lock)\n", (size_t)e.sym_idx); break;
    }
}

struct Sink {
    opvd_handle* h;
    bool quiet, raw;
    int decoded = 0, perfect = 0;

    // drain tracker lines and frames in symbol order; the tracker line of a symbol precedes its frame
    int drain() {
        std::vector<opvd_event> ev(4096);
        std::vector<uint8_t> fr(64 * OPVD_FRAME_BYTES);
        std::vector<opvd_frame_info> fi(64);
        std::vector<opvd_event> events;
        for (;;) {
            int n = opvd_poll_events(h, 0, (int)ev.size(), ev.data());
            if (n < 0) return n;
            events.insert(events.end(), ev.begin(), ev.begin() + n);
            if (n < (int)ev.size()) break;
        }
        size_t ei = 0;
        for (;;) {
            int n = opvd_poll_frames(h, (int)fi.size(), fr.data(), fi.data());
            if (n < 0) return n;
            for (int k = 0; k < n; ++k) {
                while (ei < events.size() && events[ei].sym_idx <= fi[k].ready_idx) print_event(events[ei++]);
                ++decoded;
                if (fi[k].metric == 0) ++perfect;
                const uint8_t* f = &fr[(size_t)k * OPVD_FRAME_BYTES];
                if (!quiet) print_frame(decoded, f, fi[k].metric, fi[k].sync_quality);
                if (raw) {
                    fwrite(f, 1, OPVD_FRAME_BYTES, stdout);
                    fflush(stdout);
                }
            }
            if (n < (int)fi.size()) break;
        }
        while (ei < events.size()) print_event(events[ei++]);
        return 0;
    }
};

int die(opvd_handle* h, const char* what, int rc) {
    fprintf(stderr, "opv-demod: %s: %s (%s)\n", what, opvd_strerror(rc), h ? opvd_last_cuda_error(h) : "");
    if (h) opvd_destroy(h);
    return 2;
}

bool read_full(std::vector<int16_t>& buf, size_t want_samples, size_t& got_samples) {
    // read up to want_samples I/Q pairs; returns false at EOF with got_samples possibly short
    buf.resize(want_samples * 2);
    size_t bytes = 0, want = want_samples * 4;
    char* p = reinterpret_cast<char*>(buf.data());
    while (bytes < want) {
        ssize_t n = read(STDIN_FILENO, p + bytes, want - bytes);
        if (n <= 0) break;
        bytes += (size_t)n;
    }
    got_samples = bytes / 4;  // a trailing partial sample is dropped, like cin.read (:1022)
    return bytes == want;
}

}  // namespace

int main(int argc, char* argv[]) {
    bool quiet = false, raw = false, coherent = false, streaming = false, have_init = false;
    double afc_bw = 0.001, init_offset = 0.0, pll_bw = 50.0;  // :945-947
    int device = -1;
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "-q")) quiet = true;
        else if (!strcmp(argv[i], "-r")) raw = true;
        else if (!strcmp(argv[i], "-c")) coherent = true;
        else if (!strcmp(argv[i], "-s")) streaming = true;
        else if (!strcmp(argv[i], "-a") && i + 1 < argc) afc_bw = atof(argv[++i]);
        else if (!strcmp(argv[i], "-p") && i + 1 < argc) pll_bw = atof(argv[++i]);
        else if (!strcmp(argv[i], "-o") && i + 1 < argc) { init_offset = atof(argv[++i]); have_init = true; }
        else if (!strcmp(argv[i], "--device") && i + 1 < argc) device = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-h")) {
            fprintf(stderr, "Usage: %s [options] < input.iq\n\n", argv[0]);
            fprintf(stderr, "Options:\n");
            fprintf(stderr, "  -q          Quiet mode\n");
            fprintf(stderr, "  -r          Raw output to stdout\n");
            fprintf(stderr, "  -s          Streaming mode (for live PlutoSDR input)\n");
            fprintf(stderr, "  -c          Coherent mode (Costas loop, ~3dB better)\n");
            fprintf(stderr, "  -a <bw>     AFC bandwidth (default: 0.001)\n");
            fprintf(stderr, "  -o <hz>     Initial frequency offset (streaming mode)\n");
            fprintf(stderr, "  -p <hz>     PLL bandwidth in Hz (default: 50, coherent only)\n");
            fprintf(stderr, "  --device n  CUDA device ordinal (extension)\n");
            fprintf(stderr, "  -h          Help\n");
            return 0;
        }
    }
    static char stdout_buffer[OPVD_FRAME_BYTES];
    setvbuf(stdout, stdout_buffer, _IOFBF, OPVD_FRAME_BYTES);  // one frame == one write (:978-979)

    if (!quiet) {
        fprintf(stderr, "╔═══════════════════════════════════════════════════════════════════╗\n");
        if (coherent) fprintf(stderr, "║       OPV MSK Demodulator with Costas Loop v1.0 (coherent)       ║\n");
        else if (streaming) fprintf(stderr, "║       OPV MSK Demodulator with AFC v1.0 (streaming)              ║\n");
        else fprintf(stderr, "║           OPV MSK Demodulator with AFC v1.0                       ║\n");
        fprintf(stderr, "╚═══════════════════════════════════════════════════════════════════╝\n\n");
    }

    opvd_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.n_streams = 1;
    cfg.mode = streaming ? OPVD_MODE_STREAM : OPVD_MODE_BATCH;
    cfg.afc_alpha = afc_bw;
    cfg.have_init_offset = have_init ? 1 : 0;
    cfg.init_offset_hz = init_offset;
    cfg.device = device;
    cfg.coherent = coherent ? 1 : 0;  // honoured in batch mode only, like the reference (:995 returns first)
    cfg.pll_bw_hz = pll_bw;
    opvd_handle* h = nullptr;
    int rc;

    if (streaming) {
        if (!quiet) fprintf(stderr, "Streaming mode: processing data as it arrives...\n\n");
        if (have_init && !quiet) fprintf(stderr, "Initial frequency offset: %.1f Hz\n", init_offset);
        cfg.max_samples = 8 * OPVD_CHUNK_SAMPLES;
        cfg.max_frames = 64;
        if ((rc = opvd_create(&cfg, &h)) != OPVD_OK) return die(nullptr, "create", rc);
        Sink sink{h, quiet, raw};
        std::vector<int16_t> buf;
        size_t total_in = 0, chunks = 0, symbols_in_chunks = 0;  // the EOF flush is not counted (:1067 vs :1088)
        bool eof = false, est_printed = have_init;
        opvd_stream_info si{};
        while (!eof) {
            // feed exactly up to the end of the next 86,720-sample call so that one run == one chunk (:1026)
            const size_t need_end = (size_t)si.n_samples_used + OPVD_CHUNK_SAMPLES;
            size_t want = need_end > total_in ? need_end - total_in : 0, got = 0;
            if (want) {
                eof = !read_full(buf, want, got);
                if (got && (rc = opvd_push_iq(h, 0, buf.data(), (int64_t)got)) != OPVD_OK) return die(h, "push", rc);
                total_in += got;
            }
            if ((rc = opvd_run(h, eof ? 1 : 0)) != OPVD_OK) return die(h, "run", rc);
            if ((rc = opvd_get_stream_info(h, 0, &si)) != OPVD_OK) return die(h, "info", rc);
            if (!eof) { ++chunks; symbols_in_chunks = (size_t)si.n_symbols; }
            if (!est_printed && chunks >= 1) {
                if (!quiet) fprintf(stderr, "Estimated carrier offset: %.1f Hz\n\n", si.est_offset_hz);
                est_printed = true;
            }
            if ((rc = sink.drain()) != 0) return die(h, "poll", rc);
            if (!eof && !quiet) {  // periodic status (:1079-1083)
                const size_t total_samples = chunks * (size_t)OPVD_CHUNK_SAMPLES;
                if (total_samples % (size_t)(2168000.0 * 5) < (size_t)OPVD_CHUNK_SAMPLES)
                    fprintf(stderr, "[%.1fs] %zu symbols, %d frames (%d perfect), AFC: %.1f Hz, TFreq: %.4f\n",
                            total_samples / 2168000.0, (size_t)si.n_symbols, sink.decoded, sink.perfect,
                            si.freq_offset_hz, si.timing_freq);
            }
        }
        if (!quiet) {
            const char* bar = "════════════════════════════════════════════════════════════════════";
            fprintf(stderr, "\n%s\n", bar);
            fprintf(stderr, "Summary: %d frames (%d perfect, %d errors)\n", sink.decoded, sink.perfect,
                    sink.decoded - sink.perfect);
            fprintf(stderr, "Total: %.3f sec, %zu symbols\n", chunks * (double)OPVD_CHUNK_SAMPLES / 2168000.0,
                    symbols_in_chunks);
            fprintf(stderr, "Final state: %s, AFC: %.1f Hz\n", state_name(si.sync_state), si.freq_offset_hz);
            fprintf(stderr, "%s\n", bar);
        }
        const int decoded = sink.decoded;
        opvd_destroy(h);
        return decoded > 0 ? 0 : 1;
    }

    // batch: load everything, then process (:1127-1216)
    std::vector<int16_t> all, buf;
    for (;;) {
        size_t got = 0;
        const bool more = read_full(buf, 1 << 20, got);
        all.insert(all.end(), buf.begin(), buf.begin() + got * 2);
        if (!more) break;
    }
    const size_t n = all.size() / 2;
    if (!quiet) fprintf(stderr, "Loaded %zu samples (%.3f sec)\n", n, n / 2168000.0);
    cfg.max_samples = (int64_t)std::max<size_t>(n, 64);
    if ((rc = opvd_create(&cfg, &h)) != OPVD_OK) return die(nullptr, "create", rc);
    if (n && (rc = opvd_push_iq(h, 0, all.data(), (int64_t)n)) != OPVD_OK) return die(h, "push", rc);
    if ((rc = opvd_run(h, 1)) != OPVD_OK) return die(h, "run", rc);
    opvd_stream_info si{};
    if ((rc = opvd_get_stream_info(h, 0, &si)) != OPVD_OK) return die(h, "info", rc);
    if (!quiet) {
        fprintf(stderr, "Estimated carrier offset: %.1f Hz\n", si.est_offset_hz);
        if (coherent) fprintf(stderr, "PLL bandwidth: %.1f Hz\n", pll_bw);  // :1157-1158
        fprintf(stderr, "Demodulated %zu symbols, final AFC offset: %.1f Hz\n\n", (size_t)si.n_symbols, si.freq_offset_hz);
    }
    Sink sink{h, quiet, raw};
    if ((rc = sink.drain()) != 0) return die(h, "poll", rc);
    if (!quiet) {
        const char* bar = "════════════════════════════════════════════════════════════════════";
        fprintf(stderr, "%s\n", bar);
        fprintf(stderr, "Summary: %d frames (%d perfect, %d errors)\n", sink.decoded, sink.perfect,
                sink.decoded - sink.perfect);
        fprintf(stderr, "Final state: %s, AFC: %.1f Hz\n", state_name(si.sync_state), si.freq_offset_hz);
        fprintf(stderr, "%s\n", bar);
    }
    const int decoded = sink.decoded;
    opvd_destroy(h);
    return decoded > 0 ? 0 : 1;
}
