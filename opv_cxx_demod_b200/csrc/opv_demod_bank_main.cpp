// opv-demod-bank — multi-stream, multi-GPU front-end over the same C ABI as the drop-in `opv-demod` (include/opvd.h).
//
// The reference demodulates one stdin stream per process (/root/reference/src/opv-demod.cpp:943-1217); its callers
// spawn one `opv-demod -s -r` per receiver and feed it through a pipe (src/opv-modem.cpp:391,714; `opv-modem -R`,
// :673-838, forwards every 134-byte frame as one UDP datagram, :782).  This tool sits where N of those pairs sit
// today: N inputs in, N frame streams out, all streams demodulated as ONE channel bank, sharded over the GPUs of
// the box.  Per stream the result is byte-for-byte what `opv-demod [-s] -r -q` writes to stdout for the same
// input bytes (tests/test_gpu_parity.py::test_bank_cli_*).
//
//   opv-demod-bank [-s] [-c] [-a alpha] [-p hz] [-o hz] [--devices LIST|all] [--device n] [--tile SAMPLES]
//                  [-d outdir] [-l listfile] [-u port] [--udp-in PORT --streams N [--idle-exit SEC]] [-q] INPUT...
//
// INPUT   a regular file, a FIFO or any other readable path carrying raw interleaved int16 LE I/Q (the reference's
//         stdin bytes).  FIFOs are read as they fill (live ingest): every stream advances with whatever it has
//         received, a silent stream holds nobody up.  `--udp-in PORT --streams N`: stream k instead receives its
//         samples as UDP datagrams on 127.0.0.1:(PORT + k); the tool runs until it has been idle for --idle-exit
//         seconds (default: forever).
// OUTPUT  <outdir>/<basename>.frames (default outdir: next to the input; udp<k>.frames for UDP inputs): concatenated
//         134-byte frames in stream order, flushed as they complete.  -u PORT: stream k's frames are also sent, one
//         datagram each, to 127.0.0.1:(PORT + k) — the RX egress of `opv-modem -R` for a whole bank.
// -s      streaming semantics (86,720-sample calls with carry, :1012-1113), fed in time tiles through the library's
//         sample rings: bounded host and device memory for any stream length.  Without -s: batch semantics (every
//         input is read to EOF, then one call per stream, :1127-1216).
// --devices  block-partitions the streams over the listed GPUs (stream k of n goes to the rank that owns
//         [r n/R, (r+1) n/R)); one host thread and one library handle per GPU, no data-path traffic
//         between them; the per-GPU counters are summed with one ncclAllReduce on the handles' device counters.
// Exit code 0 iff at least one frame was decoded in any stream (the per-process rule of :1124 applied to the bank).
#include <arpa/inet.h>
#include <fcntl.h>
#include <netinet/in.h>
#include <poll.h>
#include <sys/socket.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <thread>
#include <vector>

#include "../../include/opvd.h"

#ifdef OPVD_HAVE_NCCL
#include <cuda_runtime.h>
#include <nccl.h>
#endif

namespace {

struct Source {
    std::string path, out_path;
    int fd = -1;
    bool is_fifo = false, is_udp = false, seen_data = false, eof = false;
    FILE* out = nullptr;
    std::vector<uint8_t> stage;  // bytes received, not yet pushed (a trailing partial sample stays here, :1022)
    std::vector<uint8_t> all;    // batch mode: the whole input
    int decoded = 0, perfect = 0;
};

struct Options {
    bool quiet = false, coherent = false, streaming = false, have_init = false;
    double afc_bw = 0.001, init_offset = 0.0, pll_bw = 50.0;  // the reference's defaults (:945-947)
    int64_t tile = 4 * (int64_t)OPVD_CHUNK_SAMPLES;
    double idle_exit = -1.0;
    int udp_out = 0;
};

struct Shard {
    int device = -1, first = 0, count = 0, rc = 0, total = 0, perfect = 0;
    opvd_handle* h = nullptr;
    uint64_t counters[OPVD_NUM_COUNTERS] = {};
    std::string err;
};

std::string base_name(const std::string& p) {
    const size_t k = p.find_last_of('/');
    return k == std::string::npos ? p : p.substr(k + 1);
}
const char* state_name(int s) { return s == 0 ? "HUNTING" : (s == 1 ? "VERIFYING" : "LOCKED"); }

int g_udp_sock = -1;

// frames decoded since the last poll -> the streams' output files (and UDP ports)
int drain(const Options& o, Shard& sh, std::vector<Source>& src) {
    std::vector<uint8_t> fr(1024 * OPVD_FRAME_BYTES);
    std::vector<opvd_frame_info> fi(1024);
    for (;;) {
        const int n = opvd_poll_frames(sh.h, (int)fi.size(), fr.data(), fi.data());
        if (n < 0) return n;
        for (int k = 0; k < n; ++k) {
            const int g = sh.first + fi[k].stream;  // global stream index
            Source& s = src[(size_t)g];
            fwrite(&fr[(size_t)k * OPVD_FRAME_BYTES], 1, OPVD_FRAME_BYTES, s.out);
            if (g_udp_sock >= 0) {  // one frame = one datagram (src/opv-modem.cpp:782)
                sockaddr_in dst{};
                dst.sin_family = AF_INET;
                dst.sin_port = htons((uint16_t)(o.udp_out + g));
                dst.sin_addr.s_addr = inet_addr("127.0.0.1");
                sendto(g_udp_sock, &fr[(size_t)k * OPVD_FRAME_BYTES], OPVD_FRAME_BYTES, 0, (sockaddr*)&dst, sizeof(dst));
            }
            ++s.decoded;
            if (fi[k].metric == 0) ++s.perfect;
        }
        if (n && !o.streaming) continue;
        if (n < (int)fi.size()) break;
    }
    for (int k = 0; k < sh.count; ++k) fflush(src[(size_t)(sh.first + k)].out);  // frames leave as they complete (:1061)
    return 0;
}

bool fail(Shard& sh, const char* what, int rc) {
    sh.rc = rc;
    sh.err = std::string(what) + ": " + opvd_strerror(rc) + " (" + (sh.h ? opvd_last_cuda_error(sh.h) : "") + ")";
    return false;
}

// receive what the inputs of this shard have; returns false when every input is at EOF (or the bank has been idle)
bool gather(const Options& o, Shard& sh, std::vector<Source>& src, int timeout_ms, double& idle_s) {
    std::vector<pollfd> pf;
    std::vector<int> who;
    bool all_eof = true;
    for (int k = 0; k < sh.count; ++k) {
        Source& s = src[(size_t)(sh.first + k)];
        if (s.eof) continue;
        all_eof = false;
        // this stream's tile is full: push first (a datagram must fit whole, or recv would truncate it)
        if ((int64_t)s.stage.size() + (s.is_udp ? 65536 : 1) > o.tile * 4) continue;
        pf.push_back({s.fd, POLLIN, 0});
        who.push_back(sh.first + k);
    }
    if (all_eof) return false;
    const auto t0 = std::chrono::steady_clock::now();
    bool got_any = false;
    if (!pf.empty() && poll(pf.data(), pf.size(), timeout_ms) > 0) {
        for (size_t i = 0; i < pf.size(); ++i) {
            Source& s = src[(size_t)who[i]];
            if (!(pf[i].revents & (POLLIN | POLLHUP | POLLERR))) continue;
            for (;;) {  // drain the descriptor: a socket hands out one datagram per recv
                const size_t have = s.stage.size(), room = (size_t)(o.tile * 4) - have;
                if (room < (s.is_udp ? 65536u : 1u)) break;
                s.stage.resize(have + room);
                const ssize_t n = s.is_udp ? recv(s.fd, s.stage.data() + have, room, MSG_DONTWAIT) : read(s.fd, s.stage.data() + have, room);
                s.stage.resize(have + (n > 0 ? (size_t)n : 0));
                if (n > 0) { s.seen_data = true; got_any = true; continue; }
                if (n == 0 && !s.is_udp && (!s.is_fifo || s.seen_data)) s.eof = true;  // a FIFO nobody has opened yet reads 0
                else if (n < 0 && errno != EAGAIN && errno != EWOULDBLOCK && errno != EINTR) s.eof = true;
                break;
            }
        }
    }
    bool any_seen = false;  // the idle clock starts with the first byte: a bank that is waiting for its senders is not idle
    for (int k = 0; k < sh.count; ++k) any_seen = any_seen || src[(size_t)(sh.first + k)].seen_data;
    if (got_any || !any_seen) idle_s = 0.0;
    else idle_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() + 1e-4;
    if (o.idle_exit >= 0.0 && idle_s >= o.idle_exit) {  // UDP inputs have no EOF: stop after a quiet period
        for (int k = 0; k < sh.count; ++k) src[(size_t)(sh.first + k)].eof = true;
        return false;
    }
    return true;
}

void run_shard(const Options& o, Shard& sh, std::vector<Source>& src) {
    opvd_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.n_streams = sh.count;
    cfg.mode = o.streaming ? OPVD_MODE_STREAM : OPVD_MODE_BATCH;
    cfg.afc_alpha = o.afc_bw;
    cfg.have_init_offset = o.have_init ? 1 : 0;
    cfg.init_offset_hz = o.init_offset;
    cfg.device = sh.device;
    cfg.coherent = o.coherent ? 1 : 0;
    cfg.pll_bw_hz = o.pll_bw;
    int rc;
    if (!o.streaming) {
        // batch: read every input to EOF, then one call per stream over the whole capture (:1127-1216)
        int64_t max_n = 64;
        for (int k = 0; k < sh.count; ++k) {
            Source& s = src[(size_t)(sh.first + k)];
            uint8_t buf[1 << 16];
            for (;;) {
                const ssize_t n = read(s.fd, buf, sizeof(buf));
                if (n > 0) s.all.insert(s.all.end(), buf, buf + n);
                else if (n == 0 || (errno != EAGAIN && errno != EINTR)) break;
                else { pollfd p{s.fd, POLLIN, 0}; poll(&p, 1, 100); }
            }
            max_n = std::max<int64_t>(max_n, (int64_t)(s.all.size() / 4));
        }
        cfg.max_samples = max_n;
        if ((rc = opvd_create(&cfg, &sh.h)) != OPVD_OK) { fail(sh, "create", rc); return; }
        for (int k = 0; k < sh.count; ++k) {
            Source& s = src[(size_t)(sh.first + k)];
            const int64_t n = (int64_t)(s.all.size() / 4);
            if (n && (rc = opvd_push_iq(sh.h, k, reinterpret_cast<const int16_t*>(s.all.data()), n)) != OPVD_OK) { fail(sh, "push", rc); return; }
        }
        if ((rc = opvd_run(sh.h, 1)) != OPVD_OK) { fail(sh, "run", rc); return; }
        if ((rc = drain(o, sh, src)) != 0) { fail(sh, "poll", rc); return; }
    } else {
        // stream: per-stream sample rings of four tiles on the device; push what has arrived, run, drain, repeat
        cfg.max_samples = 4 * o.tile;
        cfg.max_frames = (int32_t)(2 * o.tile / OPVD_CHUNK_SAMPLES + 8);
        if ((rc = opvd_create(&cfg, &sh.h)) != OPVD_OK) { fail(sh, "create", rc); return; }
        double idle = 0.0;
        for (bool more = true; more;) {
            more = gather(o, sh, src, 20, idle);
            bool pushed = false;
            for (int k = 0; k < sh.count; ++k) {
                Source& s = src[(size_t)(sh.first + k)];
                const int64_t n = (int64_t)(s.stage.size() / 4);
                if (n <= 0) continue;
                if ((rc = opvd_push_iq(sh.h, k, reinterpret_cast<const int16_t*>(s.stage.data()), n)) != OPVD_OK) { fail(sh, "push", rc); return; }
                s.stage.erase(s.stage.begin(), s.stage.begin() + n * 4);
                pushed = true;
            }
            if (!pushed && more) continue;
            if ((rc = opvd_run(sh.h, more ? 0 : 1)) != OPVD_OK) { fail(sh, "run", rc); return; }
            if ((rc = drain(o, sh, src)) != 0) { fail(sh, "poll", rc); return; }
        }
    }
    for (int k = 0; k < sh.count; ++k) {
        sh.total += src[(size_t)(sh.first + k)].decoded;
        sh.perfect += src[(size_t)(sh.first + k)].perfect;
    }
    opvd_get_counters(sh.h, sh.counters, OPVD_NUM_COUNTERS);
}

std::vector<int> parse_devices(const std::string& s) {
    std::vector<int> d;
    if (s == "all") {
#ifdef OPVD_HAVE_NCCL
        int n = 0;
        if (cudaGetDeviceCount(&n) == cudaSuccess)
            for (int i = 0; i < n; ++i) d.push_back(i);
#endif
        return d;
    }
    size_t p = 0;
    while (p < s.size()) {
        size_t q = s.find(',', p);
        if (q == std::string::npos) q = s.size();
        const std::string tok = s.substr(p, q - p);
        const size_t dash = tok.find('-');
        if (dash != std::string::npos && dash > 0) {
            for (int i = atoi(tok.substr(0, dash).c_str()); i <= atoi(tok.substr(dash + 1).c_str()); ++i) d.push_back(i);
        } else if (!tok.empty()) d.push_back(atoi(tok.c_str()));
        p = q + 1;
    }
    return d;
}

}  // namespace

int main(int argc, char* argv[]) {
    Options o;
    std::vector<int> devices;
    int udp_in = 0, udp_streams = 0;
    std::string outdir;
    std::vector<std::string> files;
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "-q")) o.quiet = true;
        else if (!strcmp(argv[i], "-c")) o.coherent = true;
        else if (!strcmp(argv[i], "-s")) o.streaming = true;
        else if (!strcmp(argv[i], "-a") && i + 1 < argc) o.afc_bw = atof(argv[++i]);
        else if (!strcmp(argv[i], "-p") && i + 1 < argc) o.pll_bw = atof(argv[++i]);
        else if (!strcmp(argv[i], "-o") && i + 1 < argc) { o.init_offset = atof(argv[++i]); o.have_init = true; }
        else if (!strcmp(argv[i], "--device") && i + 1 < argc) devices = {atoi(argv[++i])};
        else if (!strcmp(argv[i], "--devices") && i + 1 < argc) devices = parse_devices(argv[++i]);
        else if (!strcmp(argv[i], "--tile") && i + 1 < argc) o.tile = std::max<int64_t>(atoll(argv[++i]), OPVD_CHUNK_SAMPLES);
        else if (!strcmp(argv[i], "--udp-in") && i + 1 < argc) udp_in = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--streams") && i + 1 < argc) udp_streams = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--idle-exit") && i + 1 < argc) o.idle_exit = atof(argv[++i]);
        else if (!strcmp(argv[i], "-d") && i + 1 < argc) outdir = argv[++i];
        else if (!strcmp(argv[i], "-u") && i + 1 < argc) o.udp_out = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-l") && i + 1 < argc) {
            std::ifstream lf(argv[++i]);
            for (std::string line; std::getline(lf, line);)
                if (!line.empty()) files.push_back(line);
        } else if (!strcmp(argv[i], "-h")) {
            fprintf(stderr,
                    "Usage: %s [-s] [-c] [-a bw] [-p hz] [-o hz] [--devices LIST|all] [--tile samples] [-d outdir] [-l listfile]\n"
                    "          [-u port] [--udp-in port --streams n [--idle-exit sec]] [-q] INPUT...\n"
                    "  every INPUT (file or FIFO, int16 LE I/Q) is one stream of the bank; frames go to <outdir>/<basename>.frames\n"
                    "  -s/-c/-a/-p/-o as in opv-demod; all streams share them\n"
                    "  --devices 0,1,2 | 0-7 | all: shard the streams over several GPUs (counters reduced with NCCL)\n"
                    "  -u port: also send stream k's frames as UDP datagrams to 127.0.0.1:(port + k), like opv-modem -R\n"
                    "  --udp-in port --streams n: stream k receives its samples as datagrams on 127.0.0.1:(port + k)\n",
                    argv[0]);
            return 0;
        } else files.push_back(argv[i]);
    }
    if (devices.empty()) devices = {-1};
    const size_t n_streams = udp_in > 0 ? (size_t)std::max(udp_streams, 0) : files.size();
    if (n_streams == 0) {
        fprintf(stderr, "opv-demod-bank: no inputs (-h for help)\n");
        return 2;
    }
    if (udp_in > 0 && !o.streaming) {
        fprintf(stderr, "opv-demod-bank: --udp-in needs -s (a datagram stream has no end to wait for)\n");
        return 2;
    }
    if (o.udp_out > 0) {
        if (o.udp_out + (long)n_streams > 65536 || (g_udp_sock = socket(AF_INET, SOCK_DGRAM, 0)) < 0) {
            fprintf(stderr, "opv-demod-bank: cannot open the UDP egress at port %d\n", o.udp_out);
            return 2;
        }
    }

    std::vector<Source> src(n_streams);
    for (size_t k = 0; k < n_streams; ++k) {
        Source& s = src[k];
        if (udp_in > 0) {
            s.is_udp = true;
            s.path = "udp" + std::to_string(k);
            s.fd = socket(AF_INET, SOCK_DGRAM, 0);
            sockaddr_in a{};
            a.sin_family = AF_INET;
            a.sin_port = htons((uint16_t)(udp_in + (int)k));
            a.sin_addr.s_addr = inet_addr("127.0.0.1");
            int big = 8 << 20;
            setsockopt(s.fd, SOL_SOCKET, SO_RCVBUF, &big, sizeof(big));
            if (s.fd < 0 || bind(s.fd, (sockaddr*)&a, sizeof(a)) != 0) {
                fprintf(stderr, "opv-demod-bank: cannot bind UDP port %d\n", udp_in + (int)k);
                return 2;
            }
        } else {
            s.path = files[k];
            struct stat sb;
            s.is_fifo = stat(s.path.c_str(), &sb) == 0 && S_ISFIFO(sb.st_mode);
            s.fd = open(s.path.c_str(), O_RDONLY | O_NONBLOCK);
            if (s.fd < 0) {
                fprintf(stderr, "opv-demod-bank: cannot open %s\n", s.path.c_str());
                return 2;
            }
        }
        s.out_path = (outdir.empty() ? s.path : outdir + "/" + base_name(s.path)) + ".frames";
        s.out = fopen(s.out_path.c_str(), "wb");
        if (!s.out) {
            fprintf(stderr, "opv-demod-bank: cannot create %s\n", s.out_path.c_str());
            return 2;
        }
    }

    // ---- one shard (handle + host thread) per GPU, streams block-partitioned
    const size_t R = std::min(devices.size(), n_streams);
    std::vector<Shard> shards(R);
    for (size_t r = 0; r < R; ++r) {
        const size_t base = n_streams / R, rem = n_streams % R;
        shards[r].device = devices[r];
        shards[r].first = (int)(r * base + std::min(r, rem));
        shards[r].count = (int)(base + (r < rem ? 1 : 0));
    }
    std::vector<std::thread> th;
    for (size_t r = 0; r < R; ++r) th.emplace_back([&, r] { run_shard(o, shards[r], src); });
    for (auto& t : th) t.join();
    for (size_t r = 0; r < R; ++r)
        if (shards[r].rc != 0) {
            fprintf(stderr, "opv-demod-bank: GPU %d: %s\n", shards[r].device, shards[r].err.c_str());
            return 2;
        }

    // ---- bank totals: sum of the per-GPU device counters (NCCL all-reduce when there are several GPUs)
    uint64_t totals[OPVD_NUM_COUNTERS] = {};
    const char* how = "single GPU";
#ifdef OPVD_HAVE_NCCL
    if (R > 1) {
        std::vector<ncclComm_t> comms(R);
        std::vector<int> devs(R);
        for (size_t r = 0; r < R; ++r) devs[r] = shards[r].device;
        if (ncclCommInitAll(comms.data(), (int)R, devs.data()) == ncclSuccess) {
            ncclGroupStart();
            for (size_t r = 0; r < R; ++r) {
                void* p = nullptr;
                opvd_counters_device_ptr(shards[r].h, &p);
                cudaSetDevice(devs[r]);
                ncclAllReduce(p, p, OPVD_NUM_COUNTERS, ncclUint64, ncclSum, comms[r], 0);
            }
            ncclGroupEnd();
            for (size_t r = 0; r < R; ++r) {
                cudaSetDevice(devs[r]);
                cudaStreamSynchronize(0);
            }
            opvd_get_counters(shards[0].h, totals, OPVD_NUM_COUNTERS);  // every rank now holds the bank totals
            for (auto& c : comms) ncclCommDestroy(c);
            how = "ncclAllReduce over the GPUs' device counters";
        }
    }
#endif
    if (!strcmp(how, "single GPU") || totals[OPVD_CTR_SAMPLES] == 0) {
        for (size_t r = 0; r < R; ++r)
            for (int i = 0; i < OPVD_NUM_COUNTERS; ++i) totals[i] += shards[r].counters[i];
        if (R > 1) how = "host sum (NCCL unavailable)";
    }

    int total = 0, total_perfect = 0;
    for (size_t r = 0; r < R; ++r) {
        for (int k = 0; k < shards[r].count; ++k) {
            const int g = shards[r].first + k;
            Source& s = src[(size_t)g];
            if (!o.quiet) {
                opvd_stream_info si{};
                opvd_get_stream_info(shards[r].h, k, &si);
                fprintf(stderr, "stream %d %s: %d frames (%d perfect, %d errors), %zu symbols, %s, AFC: %.1f Hz -> %s\n", g,
                        s.path.c_str(), s.decoded, s.perfect, s.decoded - s.perfect, (size_t)si.n_symbols,
                        state_name(si.sync_state), si.freq_offset_hz, s.out_path.c_str());
            }
            close(s.fd);
            fclose(s.out);
        }
        total += shards[r].total;
        total_perfect += shards[r].perfect;
        opvd_destroy(shards[r].h);
    }
    if (!o.quiet) {
        fprintf(stderr, "Summary: %zu streams, %d frames (%d perfect, %d errors)\n", n_streams, total, total_perfect,
                total - total_perfect);
        fprintf(stderr, "Bank counters (%zu GPU(s), %s): samples %llu, symbols %llu, frames decoded %llu, sync acquired %llu, "
                        "sync missed %llu, lock lost %llu\n", R, how, (unsigned long long)totals[OPVD_CTR_SAMPLES],
                (unsigned long long)totals[OPVD_CTR_SYMBOLS], (unsigned long long)totals[OPVD_CTR_FRAMES_DECODED],
                (unsigned long long)totals[OPVD_CTR_SYNC_ACQ], (unsigned long long)totals[OPVD_CTR_SYNC_MISS],
                (unsigned long long)totals[OPVD_CTR_LOST_LOCK]);
    }
    if (g_udp_sock >= 0) close(g_udp_sock);
    return total > 0 ? 0 : 1;
}
