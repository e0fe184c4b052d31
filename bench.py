#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native opv-demod receive chain.

Metric (BASELINE.json): aggregate demod Msamples/s (and decoded frames/s) over N B200s, against
the HBM roofline of the front-end kernel, with the reference's CPU opv-demod timed beside it.

Workload at N=1 = BASELINE.json configs[1]: 1,024 streams x 10 s (250 frames, 21.68 M samples,
86.7 MB each; 88.8 GB resident) of synthetic opv-mod-like captures with AWGN, Eb/N0 swept
2..10 dB across streams.  N>1: every rank owns its own 1,024 streams ("scaling": "weak").
A step = one pass of the whole chain (estimate -> demod -> sync tracker -> Viterbi) over the rank's bank
in streaming mode (`opv-demod -s` semantics), from fresh per-stream state, fed in time tiles (the
tracker + Viterbi of tile t run on a second CUDA stream while tile t+1 is demodulated).

  value   device-timed (CUDA events on the library's streams), inputs resident in HBM
  e2e     same chain through the C ABI from pinned HOST buffers: H2D of the samples, run, D2H of
          the decoded frames, on a bounded duration of the same streams (rate metric, PCIe-bound)
  channel_bank           north-star regime (BASELINE configs[4]): 18,944 streams per GPU resident, weak scaling
  channel_bank_strong    configs[4] as written: ONE 16,384-stream bank split over the N ranks (strong scaling)
  channel_bank.sustained a bank larger than HBM through push/run/poll from pinned host memory (sample/soft rings)
  --impl reference       the reference's own CPU opv-demod (oracle/_ref), one process per host core, on full
                         10-s captures of the same workload
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FS = 2168000.0
FRAME_SAMPLES = 86720
METRIC = "aggregate_demod_msps"
UNIT = "Msamples/s"
MAX_LEAD = 4000


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--streams", type=int, default=1024, help="streams per GPU")
    ap.add_argument("--seconds", type=float, default=10.0, help="capture length per stream")
    ap.add_argument("--lanes", type=int, default=0, help="demodulator kernel variant (0 = automatic)")
    ap.add_argument("--tiles", type=int, default=10, help="time tiles per step of the headline leg")
    ap.add_argument("--e2e-seconds", type=float, default=0.52, help="capture length per stream for the e2e leg")
    ap.add_argument("--e2e-tiles", type=int, default=13, help="time tiles per e2e step (H2D of a tile overlaps the kernels of the previous one; 13 = one-frame tiles: the shorter the last tile, the less kernel time trails the last copy)")
    ap.add_argument("--e2e-ring-tiles", type=int, default=2, help="device ring of the e2e leg, in tiles (+ one frame of carry)")
    ap.add_argument("--bank-streams", type=int, default=0,
                    help="streams per GPU of the channel-bank leg (0 = 4 CTAs x 32 streams per SM, 18,944 on a B200)")
    ap.add_argument("--bank-frames", type=int, default=14, help="frames per stream of the channel-bank legs")
    ap.add_argument("--bank-tiles", type=int, default=1,
                    help="time tiles per step of the channel-bank legs (1: the whole resident bank in one run; with more, the tracker "
                         "and the decoder of tile t share the SMs with the demodulator of tile t+1, which costs the demodulator "
                         "more than the overlap saves on this kernel: measured 73.0 ms with 7 tiles against 67.3 ms with 1)")
    ap.add_argument("--strong-streams", type=int, default=16384, help="total streams of the strong-scaling bank")
    ap.add_argument("--sustained-tiles", type=int, default=30, help="one-frame tiles pushed from pinned host memory in the sustained leg")
    ap.add_argument("--no-bank", action="store_true", help="skip the channel-bank legs")
    ap.add_argument("--no-sustained", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def workload_name(streams, seconds):
    return f"{streams} streams x {seconds:g} s synthetic opv-mod captures, AWGN Eb/N0 2-10 dB (BASELINE configs[1] shape)"


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, pw = [], None, set(), []
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                smax = float(r[1])
                pw.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_min_mhz": sm[0] if sm else None, "sm_max_mhz": smax,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def _traffic(table, variant, algorithmic_bytes):
    """DRAM bytes (read + write) of one launch of the dominant kernel: the ratio ncu measured on a captured launch
    (profiles/roofline_traffic.json) scaled to this launch's algorithmic bytes; None without a capture."""
    r = (table or {}).get(variant + "_dram_bytes_per_launch_per_algorithmic_byte")
    return None if r is None else int(r * algorithmic_bytes)


def load_profile_json(name):
    p = os.path.join(ROOT, "profiles", name)
    return json.load(open(p)) if os.path.exists(p) else None


# ------------------------------------------------------------------------------------------------
def cpu_reference_run(captures, threads):
    """Time the reference's CPU opv-demod (-s -r -q), one process per capture, all started together.
    captures: list of int16 [n,2] arrays.  Returns (wall_s, total_samples, total_frames, outputs, kind, threads)."""
    from oracle import oracle as ora

    binary = ora.REF_DEMOD if os.path.exists(ora.REF_DEMOD) else None
    tmp = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
    files = []
    for k, c in enumerate(captures):
        f = os.path.join(tmp, f"opvd_bench_{os.getpid()}_{k}.iq")
        c.tofile(f)
        files.append(f)
    outs = []
    try:
        if binary:
            t0 = time.perf_counter()
            procs = [subprocess.Popen([binary, "-s", "-r", "-q"], stdin=open(f, "rb"), stdout=subprocess.PIPE,
                                      stderr=subprocess.DEVNULL) for f in files]
            outs = [p.communicate()[0] for p in procs]
            wall = time.perf_counter() - t0
            kind = "reference"
        else:  # the oracle port (single C restatement per capture)
            t0 = time.perf_counter()
            outs = [ora.run(c, True, want_soft=False).frames.tobytes() for c in captures]
            wall = time.perf_counter() - t0
            kind, threads = "port", 1
    finally:
        for f in files:
            try:
                os.remove(f)
            except OSError:
                pass
    samples = sum(int(c.shape[0]) for c in captures)
    frames = sum(len(o) // 134 for o in outs)
    return wall, samples, frames, outs, kind, threads


def cpu_captures(n_caps, n_frames):
    """Full-length captures of the headline workload for the CPU arm: clean opv-mod BERT capture (TX restatement,
    bit-identical to the reference's opv-mod) + per-capture AWGN swept 2..10 dB, scaled 0.25, random lead."""
    import numpy as np

    from tools import captures as cap

    base = cap.clean_bert(n_frames).astype(np.float32) * np.float32(0.25)
    a = cap.AMP * 0.25
    out = []
    for k in range(n_caps):
        rng = np.random.default_rng(1000 + k)
        ebn0 = 2.0 + 8.0 * (k % 64) / 63.0
        sd = np.float32(np.sqrt(a * a * cap.SPS / (0.5 * 10.0 ** (ebn0 / 10.0)) / 2.0))
        lead = (k * 997) % MAX_LEAD
        x = rng.standard_normal((base.shape[0] + lead, 2), dtype=np.float32)
        x *= sd
        x[lead:] += base
        np.rint(x, out=x)
        np.clip(x, -32768, 32767, out=x)
        out.append(x.astype(np.int16))
    return out


def reference_arm(args):
    """--impl reference: the reference's own CPU implementation on all host cores, same metric/config: every process
    demodulates a FULL capture of the headline workload (args.seconds per stream), so the fixed cost of
    estimate_offset (src/opv-demod.cpp:1030-1038) weighs what it weighs in the real workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    n_frames = int(round(args.seconds * FS / FRAME_SAMPLES))
    caps = cpu_captures(cores, n_frames)
    times, samples, frames = [], 0, 0
    kind = "reference"
    for it in range(args.warmup + args.steps):
        wall, s, f, _, kind, thr = cpu_reference_run(caps, cores)
        if it >= args.warmup:
            times.append(wall)
            samples += s
            frames += f
    total = sum(times)
    value = samples / total / 1e6
    sample = (f"{cores} of the {args.streams} streams per step, each a full {args.seconds:g} s capture ({n_frames} frames), "
              f"one opv-demod -s -r -q process per core")
    line = {
        "metric": METRIC, "value": round(value, 3), "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * total / max(len(times), 1), 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.streams, args.seconds), "sample": sample},
        "frames_per_s": round(frames / total, 2),
        "cpu_baseline": {"value": round(value, 3), "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": round(value, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
def tile_ends(n, n_frames, tiles):
    """Sample counts at which the resident captures become visible: `tiles` time tiles of whole frames."""
    tiles = max(1, min(tiles, n_frames))
    return [MAX_LEAD + FRAME_SAMPLES * (n_frames * (t + 1) // tiles) for t in range(tiles - 1)] + [n]


def tiled_step(bank, ptr, stride, ends, keepalive):
    """One pass over a resident bank from fresh state, fed in time tiles; returns the accumulated device times."""
    bank.reset()
    for k, e in enumerate(ends):
        bank.attach_device_iq(ptr, stride, e, keepalive=keepalive)
        bank.run(final=(k == len(ends) - 1), sync=False)
    bank.sync()
    return bank.last_run_ms()


def run_bank_leg(args, pkg, torch, dev, local_rank, world, S, first_stream, n_frames, tiles, seed, barrier, reduce_max_ms,
                 reduce_counters, peak, mb, label):
    """A resident bank of S streams per rank (distinct memory per stream), demodulated from fresh state."""
    n = n_frames * FRAME_SAMPLES + MAX_LEAD + 4000
    stride = (n + 63) // 64 * 64
    buf = torch.empty((S, stride), dtype=torch.int32, device=dev)
    sp = pkg.make_synth(S, n_frames, stride, n, seed=seed, ebn0_lo_db=2.0, ebn0_hi_db=10.0, max_lead=MAX_LEAD,
                        first_stream=first_stream)
    pkg.synth_bank(buf.data_ptr(), sp, device=local_rank)
    bank = pkg.DemodBank(S, streaming=True, device=local_rank, lanes_per_stream=args.lanes)
    ends = tile_ends(n, n_frames, tiles)
    steps, warm = max(2, min(args.steps, 3)), 3
    per = {"estimate": 0.0, "demod": 0.0, "track": 0.0, "decode": 0.0, "total": 0.0}
    sampler = ClockSampler(local_rank)  # this leg loads every SM's FP64 pipe: the clocks it ran at belong to its numbers
    sampler.start()
    for _ in range(warm):
        tiled_step(bank, buf.data_ptr(), stride, ends, buf)
    barrier()
    for _ in range(steps):
        ms = tiled_step(bank, buf.data_ptr(), stride, ends, buf)
        for k in per:
            per[k] += ms[k]
    barrier()
    leg_clocks = sampler.stop()
    c = bank.counters()
    bank.bert_check(sp)
    ber = bank.counters()
    dev_ms = reduce_max_ms(per["total"], dev)
    demod_ms_max = reduce_max_ms(per["demod"], dev)
    tot = reduce_counters({k: c[k] for k in ("samples", "frames_decoded", "acs")}
                          | {"bit_errors": ber["bit_errors"], "frames_compared": ber["frames_compared"]}, dev)
    value = tot["samples"] * steps / (dev_ms * 1e-3) / 1e6
    demod_ms = per["demod"] / steps
    ach = c["samples"] * 4 / (demod_ms * 1e-3) / 1e9
    out = {
        "workload": f"{label}: {S} streams x {n_frames} frames per GPU ({S * n * 4 / 1e9:.1f} GB resident, distinct memory "
                    f"per stream), AWGN Eb/N0 2-10 dB, stream mode, fresh state every step, {len(ends)} time tiles per step",
        "streams_per_gpu": S, "streams_total": tot_streams(S, world, dev, reduce_counters), "frames_per_stream": n_frames,
        "steps": steps, "warmup": warm, "tiles_per_step": len(ends),
        "value": round(value, 2), "unit": UNIT, "frames_per_s": round(tot["frames_decoded"] * steps / (dev_ms * 1e-3), 1),
        "ms_per_step": round(dev_ms / steps, 3),
        "kernel_ms_per_step": {k: round(v / steps, 3) for k, v in per.items()},
        "demod_only_msps": round(tot["samples"] * steps / (demod_ms_max * 1e-3) / 1e6, 2),
        "ber": (tot["bit_errors"] / (tot["frames_compared"] * 1072.0)) if tot["frames_compared"] else None,
        "roofline": {"bound": "hbm", "kernel": bank.demod_variant(), "achieved": round(ach, 2), "peak": peak,
                     "unit": "GB/s", "frac": round(ach / peak, 5), "launch_ms": round(demod_ms / len(ends), 3),
                     "launches_per_step": len(ends), "algorithmic_bytes_per_step": int(c["samples"] * 4),
                     "traffic": _traffic(load_profile_json("roofline_traffic.json"), bank.demod_variant(), c["samples"] * 4)},
        "note": "total is the elapsed device time of the step: tracker + Viterbi of tile t overlap the demodulator of "
                "tile t+1, the estimate (first 40,000 samples of every stream) runs once in the first tile",
        "clocks": leg_clocks,
    }
    if mb:
        dfma = c["samples"] * 12.0 / (demod_ms * 1e-3)
        out["roofline"]["fp64"] = {"achieved_dfma_per_s": round(dfma, 1), "peak_dfma_per_s": mb["dfma_per_s"],
                                   "frac": round(dfma / mb["dfma_per_s"], 4),
                                   "note": "12 DFMA per sample is the Horner correlator's algorithmic count for six full "
                                           "gates; the kernel evaluates early/late gates for the dominant tone only"}
        # the honest bar: this kernel executes 16.4 FP64-pipe instructions per sample (profiles/sass_bank_r02.txt), so
        # the FP64 pipe, not HBM, is its lower ceiling: min(HBM, FP64) in bytes per second
        fp64_ceiling = mb["dfma_per_s"] / 16.4 * 4 / 1e9
        out["roofline"]["min_hbm_fp64"] = {"fp64_instr_per_sample": 16.4, "fp64_ceiling_gbs": round(fp64_ceiling, 1),
                                           "ceiling_gbs": round(min(peak, fp64_ceiling), 1),
                                           "frac": round(ach / min(peak, fp64_ceiling), 4),
                                           "note": "FP64 issue ceiling at the measured 2.2 cycles per DFMA; with the three-register "
                                                   "operand fetches ptxas leaves in this kernel (2.6 cycles per FP64 instruction, "
                                                   "tools/microbench_rf.cu) the ceiling is 15 % lower still"}
    return out, bank, buf, sp


def tot_streams(S, world, dev, reduce_counters):
    return reduce_counters({"s": S}, dev)["s"] if world > 1 else S


def run_sustained(args, pkg, torch, np, dev, local_rank, world, S, bank_buf, barrier, reduce_max_ms):
    """A bank larger than HBM end to end: S streams fed from pinned host memory in one-frame time tiles through
    push / run / poll with persistent state (sample and soft-symbol rings, no copies on the device).  The host tile
    (frame 1 of every stream, frame-aligned) is pushed over and over, so every stream is a continuous periodic
    capture; a sample of streams is checked bit for bit against the reference binary on the same bytes."""
    T = args.sustained_tiles
    try:  # 6.6 GB of pinned host memory per rank: skip the leg on every rank if any rank cannot have it
        host = torch.empty((S, FRAME_SAMPLES), dtype=torch.int32, pin_memory=True)
    except RuntimeError:
        host = None
    if reduce_max_ms(0.0 if host is not None else 1.0, dev) > 0.0:
        return {"skipped": "pinned host buffer (%.1f GB per rank) could not be allocated on every rank" % (S * FRAME_SAMPLES * 4 / 1e9)}
    host.copy_(bank_buf[:, MAX_LEAD + FRAME_SAMPLES: MAX_LEAD + 2 * FRAME_SAMPLES])
    torch.cuda.synchronize()
    sbank = pkg.DemodBank(S, streaming=True, device=local_rank, max_samples=3 * FRAME_SAMPLES + 512, max_frames=8,
                          lanes_per_stream=args.lanes)
    frames = []
    d2h = 0

    def one_pass():
        nonlocal d2h
        sbank.reset()
        got = []
        for t in range(T):
            sbank.push_iq_host_ptr(host.data_ptr(), FRAME_SAMPLES, FRAME_SAMPLES)
            sbank.run(final=(t == T - 1), sync=False)
            if t % 4 == 3 or t == T - 1:
                fr = sbank.poll_frames(wait=(t == T - 1))  # in flight: whatever the finished runs decoded, no waiting
                d2h += int(fr.data.nbytes + fr.metric.nbytes + fr.payload_start.nbytes)
                got.append(fr)
        return got

    one_pass()  # warm-up
    d2h = 0
    barrier()
    t0 = time.perf_counter()
    frames = one_pass()
    barrier()
    el = reduce_max_ms(time.perf_counter() - t0, dev)
    lost = sbank.frames_lost()
    n_frames = sum(f.data.shape[0] for f in frames)
    # device time from the first run's start (its tile has landed) to the last run's end: T - 1 tiles cross PCIe in it
    dev_ms = reduce_max_ms(sbank.last_run_ms()["total"], dev)
    out = {"workload": f"{S} streams x {T} one-frame tiles per GPU from pinned host memory ({S * FRAME_SAMPLES * 4 * T / 1e9:.0f} GB "
                       f"through a {S * 3 * FRAME_SAMPLES * 4 / 1e9:.1f} GB device ring), poll every 4 tiles",
           "value": round(world * S * FRAME_SAMPLES * T / el / 1e6, 2), "unit": UNIT, "seconds": round(el, 3),
           "value_note": "whole job (all ranks; every rank feeds its own bank from its own pinned buffer)",
           "h2d_bytes_per_gpu": int(S * FRAME_SAMPLES * 4 * T), "d2h_bytes_rank0": d2h, "frames_rank0": n_frames,
           "frames_lost": lost, "h2d_gbs": round(world * S * FRAME_SAMPLES * 4 * T / el / 1e9, 2),
           "steady_h2d_gbs": round(world * S * FRAME_SAMPLES * 4 * (T - 1) / (dev_ms * 1e-3) / 1e9, 2) if T > 1 and dev_ms > 0 else None,
           "steady_note": "tiles 2..T over the device time between the first and the last run (CUDA events): the pushes run "
                          "back to back; `seconds` (host clock) also holds the first tile's copy, the reset and the host side "
                          "of the final poll"}
    # parity on a sample of streams against the reference binary on the same (periodic) bytes
    if int(os.environ.get("RANK", "0")) == 0:
        from oracle import oracle as ora

        k = min(8, os.cpu_count() or 1)
        pick = [int(i) for i in np.linspace(0, S - 1, k)]
        caps = [np.tile(host[s].numpy().view(np.int16).reshape(-1, 2), (T, 1)) for s in pick]
        _, _, _, outs, kind, _ = cpu_reference_run(caps, k)
        mism = 0
        for s, o in zip(pick, outs):
            ref = np.frombuffer(o, np.uint8).reshape(-1, 134)
            got = np.concatenate([f.of_stream(s) for f in frames]) if frames else np.zeros((0, 134), np.uint8)
            mism += 0 if (got.shape == ref.shape and np.array_equal(got, ref)) else 1
        out["parity"] = {"streams_checked": k, "streams_mismatching": mism, "against": kind}
    sbank.close()
    del host
    return out


# ------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    import opv_cxx_demod_b200 as pkg
    from opv_cxx_demod_b200.shard import reduce_counters, reduce_max_ms, stream_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the opv-demod CUDA path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    S = args.streams                      # per rank (weak scaling)
    lo, hi = stream_range(rank, world, S * world)
    n_frames = int(round(args.seconds * FS / FRAME_SAMPLES))
    n = n_frames * FRAME_SAMPLES + MAX_LEAD + 4000
    stride = (n + 63) // 64 * 64

    # ---- synthetic bank, resident in HBM (outside every timed region)
    bank_buf = torch.empty((S, stride), dtype=torch.int32, device=dev)
    sp = pkg.make_synth(S, n_frames, stride, n, seed=20261017, ebn0_lo_db=2.0, ebn0_hi_db=10.0, max_lead=MAX_LEAD,
                        first_stream=lo)
    pkg.synth_bank(bank_buf.data_ptr(), sp, device=local_rank)
    bank = pkg.DemodBank(S, streaming=True, device=local_rank, lanes_per_stream=args.lanes)
    ends = tile_ends(n, n_frames, args.tiles)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step():
        return tiled_step(bank, bank_buf.data_ptr(), stride, ends, bank_buf)

    for _ in range(args.warmup):
        one_step()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    t_wall0 = time.perf_counter()
    per_kernel = {"estimate": 0.0, "demod": 0.0, "track": 0.0, "decode": 0.0, "total": 0.0}
    for _ in range(args.steps):
        ms = one_step()
        for k in per_kernel:
            per_kernel[k] += ms[k]
    barrier()
    wall_ms = (time.perf_counter() - t_wall0) * 1e3
    clocks = sampler.stop()
    counters = bank.counters()            # of the last step
    bank.bert_check(sp)
    ber = bank.counters()

    dev_ms = reduce_max_ms(per_kernel["total"], dev)          # max over ranks of the device-timed K steps
    tot = reduce_counters({k: counters[k] for k in ("samples", "symbols", "frames_decoded", "frames_ready",
                                                    "sync_acq", "sync_miss", "lost_lock", "acs")}
                          | {"bit_errors": ber["bit_errors"], "frames_compared": ber["frames_compared"]}, dev)
    value = tot["samples"] * args.steps / (dev_ms * 1e-3) / 1e6
    frames_per_s = tot["frames_decoded"] * args.steps / (dev_ms * 1e-3)

    # ---- roofline of the dominant kernel (demod): algorithmic bytes = 4 B/sample (+134 B/frame, negligible)
    peak, peak_src = measured_peaks()
    demod_ms = per_kernel["demod"] / args.steps
    ach = counters["samples"] * 4 / (demod_ms * 1e-3) / 1e9
    traffic = load_profile_json("roofline_traffic.json")
    mb = load_profile_json("microbench_r01.json")
    variant = bank.demod_variant()
    roofline = {"bound": "hbm", "achieved": round(ach, 2), "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 5),
                "traffic": _traffic(traffic, variant, counters["samples"] * 4),
                "traffic_source": "ncu dram__bytes ratio of a captured launch (profiles/roofline_traffic.json) x this step's bytes",
                "kernel": variant, "launch_ms": round(demod_ms / len(ends), 3), "launches_per_step": len(ends),
                "peak_source": peak_src, "algorithmic_bytes_per_step": int(counters["samples"] * 4),
                "note": "1,024 streams are bound by the per-symbol latency of the serial timing/AFC recurrence "
                        "(one warp per stream, 6.9 warps per SM), not by HBM or a pipe; the HBM/FP64-bound regime "
                        "is the channel_bank object (DESIGN.md section 3)"}
    fp64_ops_per_sample = 12.0  # DFMAs per sample of the Horner correlator (60-sample window, 2 tones, 4 per step)
    if mb:
        roofline["fp64"] = {"achieved_dfma_per_s": round(counters["samples"] * fp64_ops_per_sample / (demod_ms * 1e-3), 1),
                            "peak_dfma_per_s": mb["dfma_per_s"],
                            "frac": round(counters["samples"] * fp64_ops_per_sample / (demod_ms * 1e-3) / mb["dfma_per_s"], 4)}
    dec_s = max(per_kernel["decode"] / args.steps * 1e-3, 1e-9)
    viterbi = {"acs_per_s": round(counters["acs"] / dec_s, 1), "decode_ms": round(per_kernel["decode"] / args.steps, 3),
               "note": "decode_ms is the elapsed time of the decoder launches; with several time tiles per step they share the "
                       "SMs with the demodulator of the next tile, so the clean ACS rate is the channel_bank leg's (one tile)"}
    if mb:
        # 15 warp-instructions per trellis step of 64 states (2 per lane) = 7.5 integer lane-operations per ACS
        viterbi["int_ops_per_acs"] = 7.5
        viterbi["int32_peak_ops_per_s"] = mb["alu_ops_per_s"]
        viterbi["frac_of_int32_peak"] = round(viterbi["acs_per_s"] * 7.5 / mb["alu_ops_per_s"], 4)
        viterbi["fp32_peak_ffma_per_s"] = mb["ffma_per_s"]
        viterbi["acs_per_ffma_slot"] = round(viterbi["acs_per_s"] / mb["ffma_per_s"], 4)
        viterbi["dpx_peak_ops_per_s"] = mb["dpx_vibmin_add_per_s"]
        viterbi["frac_of_dpx_peak"] = round(viterbi["acs_per_s"] / 2.0 / mb["dpx_vibmin_add_per_s"], 4)  # 2 states per DPX op

    # ---- e2e: host buffers through the C ABI (H2D + run + D2H of frames), bounded duration
    e2e = None
    if not args.no_e2e:
        nf_e = max(2, int(round(args.e2e_seconds * FS / FRAME_SAMPLES)))
        n_e = nf_e * FRAME_SAMPLES + MAX_LEAD
        host = torch.empty((S, n_e), dtype=torch.int32, pin_memory=True)
        host.copy_(bank_buf[:, :n_e])
        torch.cuda.synchronize()
        tiles = max(1, min(args.e2e_tiles, nf_e))
        cuts = [n_e * t // tiles // 64 * 64 for t in range(tiles)] + [n_e]
        tile_max = max(cuts[t + 1] - cuts[t] for t in range(tiles))
        # device ring of two tiles + one chunk of carry: tile t+1 crosses PCIe while the kernels of tile t run
        ebank = pkg.DemodBank(S, streaming=True, device=local_rank, max_samples=args.e2e_ring_tiles * tile_max + FRAME_SAMPLES + 256,
                              lanes_per_stream=args.lanes)
        d2h = 0

        def e2e_step():
            ebank.reset()
            for t in range(tiles):
                ebank.push_iq_host_ptr(host.data_ptr() + 4 * cuts[t], cuts[t + 1] - cuts[t], n_e)
                ebank.run(final=(t == tiles - 1), sync=False)
            return ebank.poll_frames()

        for _ in range(max(1, min(args.warmup, 2))):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            fr = e2e_step()
            d2h = int(fr.data.nbytes + fr.metric.nbytes + fr.payload_start.nbytes)
        barrier()
        e_s = time.perf_counter() - t0
        e_s = reduce_max_ms(e_s, dev)
        e_val = S * world * n_e * args.steps / e_s / 1e6
        # the platform's host-to-device ceiling with all ranks copying at once (plain pinned cudaMemcpyAsync of the same
        # buffer, no kernels): what the e2e rate is a fraction of
        dst = torch.empty((S, n_e), dtype=torch.int32, device=dev)
        dst.copy_(host, non_blocking=True)
        barrier()
        t0 = time.perf_counter()
        reps_c = 3
        for _ in range(reps_c):
            dst.copy_(host, non_blocking=True)
        barrier()
        c_s = reduce_max_ms(time.perf_counter() - t0, dev)
        h2d_ceiling = S * n_e * 4 * reps_c * world / c_s / 1e9
        del dst
        # the same tile-shaped pushes through the ABI into a buffer that holds the whole capture, with no run between
        # them: what the copies alone achieve in the shape the e2e leg issues them (1,024 rows per tile, pitched)
        cbank = pkg.DemodBank(S, streaming=True, device=local_rank, max_samples=n_e + 256, lanes_per_stream=args.lanes)

        def push_only():
            cbank.reset()
            for t in range(tiles):
                cbank.push_iq_host_ptr(host.data_ptr() + 4 * cuts[t], cuts[t + 1] - cuts[t], n_e)
            cbank.sync()

        push_only()
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps_c):
            push_only()
        barrier()
        p_s = reduce_max_ms(time.perf_counter() - t0, dev)
        h2d_tiled = S * n_e * 4 * reps_c * world / p_s / 1e9
        cbank.close()
        e2e = {"value": round(e_val, 2), "unit": UNIT, "h2d_bytes_per_step": int(S * n_e * 4), "d2h_bytes_per_step": d2h,
               "sample": f"{S} streams x {nf_e} frames ({n_e} samples) per rank per step from pinned host memory, "
                         f"pushed and run in {tiles} time tiles through a {args.e2e_ring_tiles * tile_max + FRAME_SAMPLES + 256}-sample device ring "
                         f"per stream; the rate is bound by the host-to-device copies ({S * n_e * 4 * args.steps / e_s / 1e9:.1f} GB/s "
                         f"per GPU), so it does not depend on the capture length",
               "frames_per_step": int(fr.data.shape[0]),
               "h2d_gbs": round(S * world * n_e * 4 * args.steps / e_s / 1e9, 2),
               "h2d_ceiling_gbs": round(h2d_ceiling, 2),
               "frac_of_h2d_ceiling": round(S * world * n_e * 4 * args.steps / e_s / 1e9 / h2d_ceiling, 4),
               "h2d_tiled_push_gbs": round(h2d_tiled, 2),
               "h2d_tiled_push_note": "the same opvd_push_iq_all calls with no run between them (copies only)",
               "h2d_ceiling_note": "all ranks copying the same pinned buffers at once with plain cudaMemcpyAsync, no kernels; "
                                   "on this box every GPU hangs off NUMA node 0, so the ceiling itself does not scale "
                                   "linearly with the rank count"}
        ebank.close()
        del host

    # ---- CPU baseline beside it (rank 0, N=1 only): the reference binary, one process per core, FULL streams
    cpu_baseline, spot = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = min(os.cpu_count() or 1, S)
        hostc = bank_buf[:cores, :n].cpu().numpy().view(np.int16).reshape(cores, n, 2)
        caps = [np.ascontiguousarray(hostc[k]) for k in range(cores)]
        wall, s_c, f_c, outs, kind, thr = cpu_reference_run(caps, cores)
        cpu_baseline = {"value": round(s_c / wall / 1e6, 3), "unit": UNIT, "cores": thr, "kind": kind,
                        "frames_per_s": round(f_c / wall, 2),
                        "sample": f"streams 0..{cores - 1} of the same bank in full ({n_frames} frames, {n} samples each), "
                                  f"one opv-demod -s -r -q per core"}
        # parity spot check on the same bytes: every frame of every checked stream, whole-stream compare
        fr = bank.poll_frames()
        mism, compared, bad_streams = 0, 0, 0
        for k, o in enumerate(outs):
            ref = np.frombuffer(o, np.uint8).reshape(-1, 134)
            got = fr.of_stream(k)
            compared += ref.shape[0]
            if got.shape != ref.shape:
                bad_streams += 1
                mism += abs(got.shape[0] - ref.shape[0])
            else:
                d = int((ref != got).any(axis=1).sum())
                mism += d
                bad_streams += 1 if d else 0
        spot = {"streams": cores, "frames_compared": compared, "frame_mismatches": mism, "streams_mismatching": bad_streams,
                "compare": "whole streams (frame count and every byte)"}

    # ---- channel banks (north-star target: over 16,384 concurrent streams): same chain, same timing rules
    channel_bank = channel_bank_strong = None
    if not args.no_bank:
        bank.close()
        bank = None
        del bank_buf
        torch.cuda.empty_cache()
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        Sb = args.bank_streams or 4 * 32 * sms          # 4 resident CTAs x 32 streams on every SM: 18,944 on a B200
        channel_bank, bbank, bbuf, _ = run_bank_leg(args, pkg, torch, dev, local_rank, world, Sb, rank * Sb, args.bank_frames,
                                                    args.bank_tiles, 20261018, barrier, reduce_max_ms, reduce_counters, peak,
                                                    mb, "weak-scaled bank (per-GPU work fixed)")
        channel_bank["scaling"] = "weak"
        if not args.no_sustained:
            bbank.close()
            channel_bank["sustained"] = run_sustained(args, pkg, torch, np, dev, local_rank, world, Sb, bbuf, barrier, reduce_max_ms)
        else:
            bbank.close()
        del bbuf
        torch.cuda.empty_cache()
        # BASELINE configs[4] as written: ONE bank of 16,384 streams split over the ranks
        slo, shi = stream_range(rank, world, args.strong_streams)
        channel_bank_strong, sbank, sbuf, _ = run_bank_leg(args, pkg, torch, dev, local_rank, world, shi - slo, slo,
                                                           args.bank_frames, args.bank_tiles, 20261019, barrier, reduce_max_ms,
                                                           reduce_counters, peak, mb,
                                                           f"fixed {args.strong_streams}-stream bank split over {world} GPU(s)")
        channel_bank_strong["scaling"] = "strong"
        sbank.close()
        del sbuf
        torch.cuda.empty_cache()

    if channel_bank and mb:  # the Viterbi kernel alone (the bank leg runs the chain without overlap)
        cb = channel_bank
        acs_s = cb["frames_per_s"] * 68608.0 * cb["ms_per_step"] / max(cb["kernel_ms_per_step"]["decode"], 1e-9) if cb["tiles_per_step"] == 1 else None
        if acs_s:
            viterbi.update({"acs_per_s_alone": round(acs_s, 1), "frac_of_int32_peak_alone": round(acs_s * 7.5 / mb["alu_ops_per_s"], 4),
                            "frac_of_dpx_peak_alone": round(acs_s / 2.0 / mb["dpx_vibmin_add_per_s"], 4)})
    if rank == 0:
        n_launch = (4 * len(ends) + 1) * args.steps  # est + demod + track + decode (+ log advance) per tile
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(dev_ms / args.steps, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(S, args.seconds), "streams_per_gpu": S, "streams_total": S * world,
                       "frames_per_stream": n_frames, "samples_per_stream": n, "mode": "stream (-s)",
                       "tiles_per_step": len(ends),
                       "l2": "inputs (%.1f GB per GPU) far larger than the 126 MB L2; no flush needed" % (S * n * 4 / 1e9),
                       "lanes_per_stream": args.lanes},
            "frames_per_s": round(frames_per_s, 1),
            "wall_ms_per_step": round(wall_ms / args.steps, 3),
            "kernel_ms_per_step": {k: round(v / args.steps, 3) for k, v in per_kernel.items()},
            "counters": tot,
            "ber": (tot["bit_errors"] / (tot["frames_compared"] * 1072.0)) if tot["frames_compared"] else None,
            "roofline": roofline, "viterbi": viterbi, "cpu_baseline": cpu_baseline, "parity_spotcheck": spot,
            "channel_bank": channel_bank, "channel_bank_strong": channel_bank_strong,
            "e2e": e2e, "gpu_launches": n_launch, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if bank is not None:
        bank.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
