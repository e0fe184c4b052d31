"""tools/captures.py — seeded synthetic OPV captures for tests and the CPU-baseline leg.

TEST/MEASUREMENT TOOLING (uses oracle/ for the TX chain); not imported by the product package.

A capture = int16 little-endian interleaved I/Q, exactly what opv-demod reads on stdin
(src/opv-demod.cpp:68,1022).  Clean captures come from the TX restatement in oracle/opv_oracle.c
(bit-identical to the reference opv-mod, checked in tests/test_oracle_vs_ref.py) and are then
impaired as SURVEY.md §8(d) describes.
"""
from __future__ import annotations

import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as ora  # noqa: E402

FS = 2168000.0
SPS = 40
FRAME_SAMPLES = 2168 * SPS
AMP = 16383.0


def clean_bert(n_frames: int, callsign: str = "W5NYV", first: int = 0) -> np.ndarray:
    """opv-mod -S <callsign> -B <n> (src/opv-mod.cpp:502-529) -> int16 [N,2]."""
    return ora.modulate(ora.bert_frames(callsign, n_frames, first=first))


def clean_random(n_frames: int, seed: int) -> tuple[np.ndarray, np.ndarray]:
    """opv-mod -R on seeded random 134-byte frames. Returns (iq int16 [N,2], frames uint8 [n,134])."""
    rng = np.random.default_rng(seed)
    frames = rng.integers(0, 256, size=(n_frames, 134), dtype=np.uint8)
    return ora.modulate(frames), frames


def impair(iq: np.ndarray, seed: int, *, scale: float = 0.25, ebn0_db: float | None = None,
           cfo_hz: float = 0.0, frac_delay: float = 0.0, lead_gap: int = 0, lead_noise: bool = True,
           dropouts: list[tuple[int, int]] | None = None, tail_gap: int = 0) -> np.ndarray:
    """Apply headroom scaling, fractional delay, CFO, AWGN, leading/trailing gaps and dropouts; re-quantise.

    Eb/N0 follows SURVEY.md §8(d): total per-sample complex noise variance
    sigma^2 = A^2 * SPS / (R * 10^(EbN0/10)) with R = 1/2 and A the scaled amplitude.
    dropouts = [(start_sample, length)] spans (relative to the signal start) replaced by noise only.
    """
    rng = np.random.default_rng(seed)
    x = iq.astype(np.float64)
    z = (x[:, 0] + 1j * x[:, 1]) * scale
    if frac_delay:
        z = (1.0 - frac_delay) * z + frac_delay * np.concatenate([z[:1], z[:-1]])
    if dropouts:
        for s, ln in dropouts:
            z[s:s + ln] = 0.0
    if lead_gap or tail_gap:
        z = np.concatenate([np.zeros(lead_gap, complex), z, np.zeros(tail_gap, complex)])
    n = np.arange(z.size, dtype=np.float64)
    if cfo_hz:
        z = z * np.exp(2j * np.pi * cfo_hz * n / FS)
    if ebn0_db is not None:
        a = AMP * scale
        sigma2 = a * a * SPS / (0.5 * 10.0 ** (ebn0_db / 10.0))
        noise = rng.normal(0.0, np.sqrt(sigma2 / 2.0), size=(z.size, 2))
        if not lead_noise and lead_gap:
            noise[:lead_gap] = 0.0
        z = z + noise[:, 0] + 1j * noise[:, 1]
    out = np.empty((z.size, 2), np.float64)
    out[:, 0] = z.real
    out[:, 1] = z.imag
    return np.clip(np.rint(out), -32768, 32767).astype(np.int16)


def standard_cases():
    """Small named parity cases shared by the CPU and GPU tests: name -> (iq, description)."""
    cases = {}
    cases["clean5"] = clean_bert(5)
    cases["clean12_call"] = clean_bert(12, "KB5MU", first=250)
    base = clean_bert(8)
    cases["awgn14"] = impair(base, 101, ebn0_db=14.0)
    cases["awgn8"] = impair(base, 102, ebn0_db=8.0)
    cases["awgn4"] = impair(base, 103, ebn0_db=4.0)
    cases["cfo_p1200_delay"] = impair(base, 104, ebn0_db=16.0, cfo_hz=1200.0, frac_delay=0.37, lead_gap=12345)
    cases["cfo_m1900"] = impair(base, 105, ebn0_db=18.0, cfo_hz=-1900.0, frac_delay=0.81, lead_gap=777)
    rnd, _ = clean_random(6, 7)
    cases["random6"] = impair(rnd, 106, ebn0_db=20.0, lead_gap=40001)
    long = clean_bert(22)
    # short dropout (3 frames: flywheel) and long dropout (7 frames: lock loss + re-acquire)
    cases["dropout_short"] = impair(long, 107, ebn0_db=15.0, lead_gap=5000,
                                    dropouts=[(5 * FRAME_SAMPLES + 1000, 3 * FRAME_SAMPLES)])
    cases["dropout_long"] = impair(long, 108, ebn0_db=15.0, lead_gap=2500,
                                   dropouts=[(4 * FRAME_SAMPLES + 300, 7 * FRAME_SAMPLES)])
    cases["zeros_gap"] = impair(long, 109, ebn0_db=None, lead_gap=0,
                                dropouts=[(6 * FRAME_SAMPLES, 8 * FRAME_SAMPLES)])
    cases["noise_only"] = impair(np.zeros((3 * FRAME_SAMPLES, 2), np.int16), 110, ebn0_db=6.0)
    cases["short_lt_chunk"] = clean_bert(5)[:60000]
    cases["tiny"] = clean_bert(1)[:37]
    cases["empty"] = np.zeros((0, 2), np.int16)
    return cases


if __name__ == "__main__":
    import argparse

    ap = argparse.ArgumentParser(description="write a seeded synthetic capture (int16 LE I/Q) to stdout")
    ap.add_argument("--frames", type=int, default=10)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--ebn0", type=float, default=None)
    ap.add_argument("--cfo", type=float, default=0.0)
    ap.add_argument("--delay", type=float, default=0.0)
    ap.add_argument("--gap", type=int, default=0)
    a = ap.parse_args()
    cap = impair(clean_bert(a.frames), a.seed, ebn0_db=a.ebn0, cfo_hz=a.cfo, frac_delay=a.delay, lead_gap=a.gap)
    sys.stdout.buffer.write(cap.tobytes())
