"""Quick throughput probe on a resident synthetic bank (development aid; bench.py is the contract)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import opv_cxx_demod_b200 as pkg

ap = argparse.ArgumentParser()
ap.add_argument("--streams", type=int, default=1024)
ap.add_argument("--frames", type=int, default=25)
ap.add_argument("--mode", default="stream")
ap.add_argument("--lanes", type=int, default=0)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--tiles", type=int, default=1, help="time tiles per pass (attach a growing prefix, run each)")
ap.add_argument("--tile-frames", default="", help="comma-separated frame counts at which the time tiles end (overrides --tiles)")
ap.add_argument("--stride-frames", type=int, default=0, help="row pitch as if the captures were this many frames long (address span without more data)")
ap.add_argument("--ebn0", type=float, nargs=2, default=[2.0, 10.0])
a = ap.parse_args()
S, nf = a.streams, a.frames
n = nf * 86720 + 8000
stride = (max(n, a.stride_frames * 86720 + 8000) + 63) // 64 * 64
buf = torch.empty((S, stride), dtype=torch.int32, device="cuda")
sp = pkg.make_synth(S, nf, stride, n, seed=1, ebn0_lo_db=a.ebn0[0], ebn0_hi_db=a.ebn0[1], cfo_max_hz=0.0, max_lead=4000)
t0 = time.time(); pkg.synth_bank(buf.data_ptr(), sp); torch.cuda.synchronize(); t1 = time.time()
print(f"synth {S}x{n} samples ({S*n*4/1e9:.2f} GB) in {t1-t0:.2f}s", flush=True)
for rep in range(a.reps):
    bank = pkg.DemodBank(S, streaming=(a.mode == "stream"), lanes_per_stream=a.lanes)
    torch.cuda.synchronize()
    t0 = time.time()
    if a.tile_frames:
        ends = [4000 + 86720 * int(x) for x in a.tile_frames.split(",")] + [n]
        for k, e in enumerate(ends):
            bank.attach_device_iq(buf.data_ptr(), stride, e, keepalive=buf)
            bank.run(final=(k == len(ends) - 1), sync=False)
    elif a.tiles <= 1:
        bank.attach_device_iq(buf.data_ptr(), stride, n, keepalive=buf)
        bank.run(final=True)
    else:
        for t in range(a.tiles):
            avail = n if t == a.tiles - 1 else (t + 1) * n // a.tiles
            bank.attach_device_iq(buf.data_ptr(), stride, avail, keepalive=buf)
            bank.run(final=(t == a.tiles - 1), sync=False)
    ms = bank.last_run_ms(); c = bank.counters()
    if os.environ.get("OPVD_TRACE"):
        bank.poll_frames()  # prints the timeline of the runs (at most 8)
    t1 = time.time()
    tot = ms["total"] / 1e3
    print(json.dumps({"rep": rep, "S": S, "frames": nf, "ms": ms, "wall_s": round(t1 - t0, 4),
                      "Msps": round(S * n / tot / 1e6, 1), "GBps": round(S * n * 4 / tot / 1e9, 1),
                      "frames_per_s": round(c["frames_decoded"] / tot, 1), "counters": c}), flush=True)
    bank.close()
