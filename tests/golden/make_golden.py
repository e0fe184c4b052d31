"""Generate tests/golden/reference_golden.json from the REFERENCE ITSELF (oracle/_ref, built from
/root/reference/src by oracle/Makefile).  Run in the authoring container:  python tests/golden/make_golden.py

For every standard capture (tools/captures.standard_cases) and both modes it records what the
unmodified reference opv-demod binary produced: sha256 of the frame bytes on stdout, the parsed
tracker events from stderr, the exit code, the Summary line, and — through the #include harness —
the number of soft symbols, a sha256 of their IEEE-754 bytes and the estimated offset.
"""
import hashlib
import json
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np  # noqa: E402

from oracle import oracle as ora  # noqa: E402
from tools import captures as cap  # noqa: E402


def main():
    assert ora.have_ref(), "oracle/_ref is missing (needs /root/reference)"
    out = {"_about": "outputs of the unmodified reference opv-demod (see make_golden.py)", "cases": {}}
    for name, iq in cap.standard_cases().items():
        for streaming in (False, True):
            args = ["-r"] + (["-s"] if streaming else [])
            frames, events, rc, err = ora.run_ref_binary(iq, args)
            summ = re.search(r"Summary: .*", err)
            soft, est, ff, tf, chunks = ora.ref_run_soft(iq, streaming)
            out["cases"][f"{name}/{'stream' if streaming else 'batch'}"] = {
                "capture_sha256": hashlib.sha256(np.ascontiguousarray(iq).tobytes()).hexdigest(),
                "n_samples": int(iq.shape[0]),
                "n_frames": int(frames.shape[0]),
                "frames_sha256": hashlib.sha256(frames.tobytes()).hexdigest(),
                "events": [[int(t), int(i), int(c)] for (t, i, c) in events],
                "exit_code": int(rc),
                "summary": summ.group(0) if summ else None,
                "n_soft": int(soft.size),
                "soft_sha256": hashlib.sha256(soft.tobytes()).hexdigest(),
                "est_offset": float(est),
                "final_freq": float(ff),
                "chunk_starts": [int(c) for c in chunks],
            }
        # -c coherent mode (batch only): reference binary + CoherentMSKDemodulator through the harness
        frames, events, rc, err = ora.run_ref_binary(iq, ["-c", "-r"])
        summ = re.search(r"Summary: .*", err)
        soft, est, ff = ora.ref_run_soft_coherent(iq)
        out["cases"][f"{name}/coherent"] = {
            "capture_sha256": hashlib.sha256(np.ascontiguousarray(iq).tobytes()).hexdigest(),
            "n_samples": int(iq.shape[0]),
            "n_frames": int(frames.shape[0]),
            "frames_sha256": hashlib.sha256(frames.tobytes()).hexdigest(),
            "events": [[int(t), int(i), int(c)] for (t, i, c) in events],
            "exit_code": int(rc),
            "summary": summ.group(0) if summ else None,
            "n_soft": int(soft.size),
            "soft_sha256": hashlib.sha256(soft.tobytes()).hexdigest(),
            "est_offset": float(est),
            "final_freq": float(ff),
            "chunk_starts": [0],
        }
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_golden.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote", path, len(out["cases"]), "cases")


if __name__ == "__main__":
    main()


def stage_vectors():
    """Stage-level vectors from the reference's own FrameDecoder / ViterbiDecoder (via the harness)."""
    import ctypes as C

    R = ora.ref()
    rng = np.random.default_rng(2024)
    cases = cap.standard_cases()
    payloads = []
    for name in ("awgn4", "awgn8", "clean5", "noise_only"):
        soft, *_ = ora.ref_run_soft(cases[name], False)
        for k in range(3):
            start = 24 + k * 2168 + int(rng.integers(0, 50))
            if start + 2144 <= soft.size:
                payloads.append(soft[start:start + 2144])
    payloads.append(rng.normal(0, 1e9, 2144))          # pure noise
    payloads.append(np.zeros(2144))                    # dropped frame (scale < 1e-10)
    payloads.append(np.full(2144, 1e-12))              # dropped frame, non-zero
    tie = rng.choice([-1.0, 1.0], 2144) * 1e6          # every symbol at the same magnitude: many metric ties
    payloads.append(tie)
    payloads = np.array(payloads)
    frames = np.zeros((len(payloads), 134), np.uint8)
    metrics = np.zeros(len(payloads), np.int32)
    for i, p in enumerate(payloads):
        p = np.ascontiguousarray(p)
        metrics[i] = R.ref_frame_decode(p.ctypes.data, frames[i].ctypes.data)
    q = rng.integers(0, 8, size=(6, 2144)).astype(np.int32)
    q[1] = 3                                           # all-equal inputs: pure tie-breaking
    q[2] = rng.integers(3, 5, 2144)
    bits = np.zeros((6, 1072), np.uint8)
    vmet = np.zeros(6, np.int32)
    for i in range(6):
        vmet[i] = R.ref_viterbi_decode(q[i].ctypes.data, bits[i].ctypes.data)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stage_vectors.npz")
    np.savez_compressed(path, payloads=payloads, frames=frames, metrics=metrics, vit_in=q, vit_bits=bits, vit_metric=vmet)
    print("wrote", path, payloads.shape, metrics.tolist(), vmet.tolist())


if __name__ == "__main__":
    stage_vectors()
