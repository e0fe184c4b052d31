// opv-demod-bank — multi-stream front-end over the same C ABI as the drop-in `opv-demod` (include/opvd.h).
//
// The reference demodulates one stdin stream per process (/root/reference/src/opv-demod.cpp:943-1217; its
// callers spawn one `opv-demod -s -r` per receiver, src/opv-modem.cpp:391,714).  This tool is the extension
// SURVEY.md section 8(b) leaves room for: N capture files in, N frame files out, all streams demodulated as
// ONE bank on one GPU.  Per stream the result is byte-for-byte what `opv-demod [-s] -r -q < FILE` writes to
// stdout (tests/test_gpu_parity.py::test_bank_cli_matches_per_stream_reference).
//
//   opv-demod-bank [-s] [-c] [-a alpha] [-p hz] [-o hz] [--device n] [-d outdir] [-l listfile] [-u port] [-q] FILE...
//
// -u PORT: the RX egress of `opv-modem -R` for a whole bank (src/opv-modem.cpp:673-838: every 134-byte frame
// read from the demodulator is one UDP datagram to 127.0.0.1:<response port>, :782): stream k's frames are
// also sent, one datagram each, to 127.0.0.1:(PORT + k), so one Interlocutor per receiver can listen on its
// own port.
//
// FILE: raw interleaved int16 LE I/Q (the reference's stdin bytes).  Output: <outdir>/<basename>.frames
// (default outdir: next to the input), concatenated 134-byte frames in stream order.  -s: streaming
// semantics (86,720-sample calls with carry, :1012-1113), fed in time tiles so that host memory stays
// bounded and frames leave as they complete; without -s: batch semantics (whole capture in one call,
// :1127-1216).  GPU memory holds the captures (4 bytes per sample and stream).  Exit code 0 iff at least one
// frame was decoded in any stream (the per-process rule of :1124 applied to the bank).
#include <arpa/inet.h>
#include <netinet/in.h>
#include <sys/socket.h>
#include <unistd.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "../../include/opvd.h"

namespace {

struct Input {
    std::string path, out_path;
    FILE* in = nullptr;
    FILE* out = nullptr;
    int64_t n_samples = 0, pushed = 0;
    int decoded = 0, perfect = 0;
};

int die(opvd_handle* h, const char* what, int rc) {
    fprintf(stderr, "opv-demod-bank: %s: %s (%s)\n", what, opvd_strerror(rc), h ? opvd_last_cuda_error(h) : "");
    if (h) opvd_destroy(h);
    return 2;
}

std::string base_name(const std::string& p) {
    const size_t k = p.find_last_of('/');
    return k == std::string::npos ? p : p.substr(k + 1);
}

const char* state_name(int s) { return s == 0 ? "HUNTING" : (s == 1 ? "VERIFYING" : "LOCKED"); }

int g_udp_sock = -1, g_udp_port = 0;

// frames decoded so far -> the streams' output files (and UDP ports), in (stream, frame) order as the library
// returns them
int drain(opvd_handle* h, std::vector<Input>& in) {
    std::vector<uint8_t> fr(256 * OPVD_FRAME_BYTES);
    std::vector<opvd_frame_info> fi(256);
    for (;;) {
        const int n = opvd_poll_frames(h, (int)fi.size(), fr.data(), fi.data());
        if (n < 0) return n;
        for (int k = 0; k < n; ++k) {
            Input& s = in[(size_t)fi[k].stream];
            fwrite(&fr[(size_t)k * OPVD_FRAME_BYTES], 1, OPVD_FRAME_BYTES, s.out);
            if (g_udp_sock >= 0) {  // one frame = one datagram (src/opv-modem.cpp:782)
                sockaddr_in dst{};
                dst.sin_family = AF_INET;
                dst.sin_port = htons((uint16_t)(g_udp_port + fi[k].stream));
                dst.sin_addr.s_addr = inet_addr("127.0.0.1");
                sendto(g_udp_sock, &fr[(size_t)k * OPVD_FRAME_BYTES], OPVD_FRAME_BYTES, 0, (sockaddr*)&dst, sizeof(dst));
            }
            ++s.decoded;
            if (fi[k].metric == 0) ++s.perfect;
        }
        if (n < (int)fi.size()) return 0;
    }
}

}  // namespace

int main(int argc, char* argv[]) {
    bool quiet = false, coherent = false, streaming = false, have_init = false;
    double afc_bw = 0.001, init_offset = 0.0, pll_bw = 50.0;  // the reference's defaults (:945-947)
    int device = -1;
    std::string outdir;
    std::vector<std::string> files;
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "-q")) quiet = true;
        else if (!strcmp(argv[i], "-c")) coherent = true;
        else if (!strcmp(argv[i], "-s")) streaming = true;
        else if (!strcmp(argv[i], "-a") && i + 1 < argc) afc_bw = atof(argv[++i]);
        else if (!strcmp(argv[i], "-p") && i + 1 < argc) pll_bw = atof(argv[++i]);
        else if (!strcmp(argv[i], "-o") && i + 1 < argc) { init_offset = atof(argv[++i]); have_init = true; }
        else if (!strcmp(argv[i], "--device") && i + 1 < argc) device = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-d") && i + 1 < argc) outdir = argv[++i];
        else if (!strcmp(argv[i], "-u") && i + 1 < argc) g_udp_port = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-l") && i + 1 < argc) {
            std::ifstream lf(argv[++i]);
            for (std::string line; std::getline(lf, line);)
                if (!line.empty()) files.push_back(line);
        } else if (!strcmp(argv[i], "-h")) {
            fprintf(stderr,
                    "Usage: %s [-s] [-c] [-a bw] [-p hz] [-o hz] [--device n] [-d outdir] [-l listfile] [-u port] [-q] FILE...\n"
                    "  every FILE (int16 LE I/Q) is one stream of the bank; frames go to <outdir>/<basename>.frames\n"
                    "  -s/-c/-a/-p/-o as in opv-demod; all streams share them\n"
                    "  -u port: also send stream k's frames as UDP datagrams to 127.0.0.1:(port + k), like opv-modem -R\n",
                    argv[0]);
            return 0;
        } else files.push_back(argv[i]);
    }
    if (files.empty()) {
        fprintf(stderr, "opv-demod-bank: no input files (-h for help)\n");
        return 2;
    }
    if (g_udp_port > 0) {
        if (g_udp_port + (long)files.size() > 65536 || (g_udp_sock = socket(AF_INET, SOCK_DGRAM, 0)) < 0) {
            fprintf(stderr, "opv-demod-bank: cannot open the UDP egress at port %d\n", g_udp_port);
            return 2;
        }
    }

    std::vector<Input> in(files.size());
    int64_t max_n = 0;
    for (size_t k = 0; k < files.size(); ++k) {
        Input& s = in[k];
        s.path = files[k];
        s.in = fopen(s.path.c_str(), "rb");
        if (!s.in) {
            fprintf(stderr, "opv-demod-bank: cannot open %s\n", s.path.c_str());
            return 2;
        }
        fseek(s.in, 0, SEEK_END);
        s.n_samples = (int64_t)(ftell(s.in) / 4);  // a trailing partial sample is dropped, like cin.read (:1022)
        fseek(s.in, 0, SEEK_SET);
        max_n = std::max(max_n, s.n_samples);
        s.out_path = (outdir.empty() ? s.path : outdir + "/" + base_name(s.path)) + ".frames";
        s.out = fopen(s.out_path.c_str(), "wb");
        if (!s.out) {
            fprintf(stderr, "opv-demod-bank: cannot create %s\n", s.out_path.c_str());
            return 2;
        }
    }

    opvd_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.n_streams = (int32_t)in.size();
    cfg.mode = streaming ? OPVD_MODE_STREAM : OPVD_MODE_BATCH;
    cfg.afc_alpha = afc_bw;
    cfg.have_init_offset = have_init ? 1 : 0;
    cfg.init_offset_hz = init_offset;
    cfg.device = device;
    cfg.coherent = coherent ? 1 : 0;
    cfg.pll_bw_hz = pll_bw;
    // stream mode: fed in tiles of 8 calls per stream and run (bounded host memory, frames leave as they complete);
    // batch mode: the whole capture is one call (:1164-1166).  The captures stay resident on the GPU in both
    // modes: the library drops consumed samples from the front of ALL rows at once, and a bank of unequal files
    // always holds a short stream that still needs its first sample at EOF.
    const int64_t tile = streaming ? 8 * (int64_t)OPVD_CHUNK_SAMPLES : std::max<int64_t>(max_n, 64);
    cfg.max_samples = std::max<int64_t>(max_n, 64);
    cfg.max_frames = (int32_t)(tile / OPVD_CHUNK_SAMPLES + 8);
    opvd_handle* h = nullptr;
    int rc = opvd_create(&cfg, &h);
    if (rc != OPVD_OK) return die(nullptr, "create", rc);

    std::vector<int16_t> buf((size_t)std::min<int64_t>(tile, std::max<int64_t>(max_n, 1)) * 2);
    for (bool more = true; more;) {
        more = false;
        for (size_t k = 0; k < in.size(); ++k) {
            Input& s = in[k];
            const int64_t want = std::min<int64_t>(tile, s.n_samples - s.pushed);
            if (want <= 0) continue;
            const size_t got = fread(buf.data(), 4, (size_t)want, s.in);
            if (got && (rc = opvd_push_iq(h, (int32_t)k, buf.data(), (int64_t)got)) != OPVD_OK) return die(h, "push", rc);
            s.pushed += (int64_t)got;
            if ((int64_t)got < want) s.n_samples = s.pushed;  // file shrank under us: what we have is the stream
            if (s.pushed < s.n_samples) more = true;
        }
        if ((rc = opvd_run(h, more ? 0 : 1)) != OPVD_OK) return die(h, "run", rc);
        if ((rc = drain(h, in)) != 0) return die(h, "poll", rc);
    }

    int total = 0, total_perfect = 0;
    for (size_t k = 0; k < in.size(); ++k) {
        Input& s = in[k];
        fclose(s.in);
        fclose(s.out);
        total += s.decoded;
        total_perfect += s.perfect;
        if (!quiet) {
            opvd_stream_info si{};
            opvd_get_stream_info(h, (int32_t)k, &si);
            fprintf(stderr, "stream %zu %s: %d frames (%d perfect, %d errors), %zu symbols, %s, AFC: %.1f Hz -> %s\n", k,
                    s.path.c_str(), s.decoded, s.perfect, s.decoded - s.perfect, (size_t)si.n_symbols,
                    state_name(si.sync_state), si.freq_offset_hz, s.out_path.c_str());
        }
    }
    if (!quiet)
        fprintf(stderr, "Summary: %zu streams, %d frames (%d perfect, %d errors)\n", in.size(), total, total_perfect,
                total - total_perfect);
    opvd_destroy(h);
    if (g_udp_sock >= 0) close(g_udp_sock);
    return total > 0 ? 0 : 1;
}
