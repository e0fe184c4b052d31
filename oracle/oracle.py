"""oracle/oracle.py — TEST INFRASTRUCTURE ONLY.

ctypes front-end for the CPU checker:
  * ``libopv_oracle.so``  — plain-C restatement (oracle/opv_oracle.c), always available (gcc).
  * ``_ref/libref_stages.so`` / ``_ref/opv-demod`` / ``_ref/opv-mod`` — the reference itself,
    compiled from /root/reference/src by oracle/Makefile (present in the authoring container and
    shipped prebuilt to the GPU box; /root/reference is never read at run time).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (opv_cxx_demod_b200) must never do so.
"""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess
from dataclasses import dataclass, field

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
REF_DEMOD = os.path.join(REF_DIR, "opv-demod")
REF_MOD = os.path.join(REF_DIR, "opv-mod")
REF_MODEM = os.path.join(REF_DIR, "opv-modem")

SPS = 40
FRAME_BYTES = 134
FRAME_BITS = 1072
ENCODED_BITS = 2144
FRAME_SYMBOLS = 2168
CHUNK_SAMPLES = 86720
FRAME_SAMPLES = FRAME_SYMBOLS * SPS

EV_NAMES = {1: "HUNT_TO_VERIFY", 2: "VERIFY_TO_LOCKED", 3: "SYNC_OK", 4: "SYNC_MISS", 5: "LOST_LOCK"}


def build(force: bool = False) -> None:
    """Compile the C restatement (and the reference, when /root/reference is present)."""
    so = os.path.join(HERE, "libopv_oracle.so")
    src = os.path.join(HERE, "opv_oracle.c")
    need = force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src)
    have_ref_src = os.path.exists("/root/reference/src/opv-demod.cpp")
    need_ref = have_ref_src and (force or not os.path.exists(os.path.join(REF_DIR, "libref_stages.so"))
                                 or not os.path.exists(REF_DEMOD) or not os.path.exists(REF_MOD)
                                 or not os.path.exists(REF_MODEM))
    if need or need_ref:
        subprocess.run(["make", "-C", HERE] + (["-B"] if force else []), check=True,
                       stdout=subprocess.DEVNULL)


class _Event(C.Structure):
    _fields_ = [("type", C.c_int32), ("count", C.c_int32), ("sym_idx", C.c_int64),
                ("corr", C.c_double), ("raw", C.c_double)]


class _Cfg(C.Structure):
    _fields_ = [("streaming", C.c_int), ("afc_alpha", C.c_double),
                ("have_init_offset", C.c_int), ("init_offset", C.c_double),
                ("coherent", C.c_int), ("pll_bw", C.c_double)]


class _Result(C.Structure):
    _fields_ = [("frames", C.c_void_p), ("metrics", C.c_void_p), ("frame_ready_idx", C.c_void_p),
                ("cap_frames", C.c_size_t),
                ("soft", C.c_void_p), ("cap_soft", C.c_size_t),
                ("events", C.c_void_p), ("cap_events", C.c_size_t),
                ("chunk_starts", C.c_void_p), ("cap_chunks", C.c_size_t),
                ("n_frames", C.c_size_t), ("n_perfect", C.c_size_t), ("n_dropped", C.c_size_t),
                ("n_soft", C.c_size_t), ("n_events", C.c_size_t), ("n_chunks", C.c_size_t),
                ("est_offset", C.c_double), ("final_freq", C.c_double), ("final_tfreq", C.c_double),
                ("final_state", C.c_int), ("total_samples", C.c_size_t)]


class _Demod(C.Structure):
    _fields_ = [(n, C.c_double) for n in
                ("freq_offset", "phase_f1", "phase_f2", "prev1_re", "prev1_im", "prev2_re", "prev2_im",
                 "afc_alpha", "mu", "timing_freq", "alpha_timing", "beta_timing")] + [("leftover", C.c_size_t)]


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(os.path.join(HERE, "libopv_oracle.so"))
        L.ora_set_codemod_perturb.argtypes = [C.c_double]
        L.ora_estimate_offset.restype = C.c_double
        L.ora_estimate_offset.argtypes = [C.c_void_p, C.c_size_t]
        L.ora_demodulate.restype = C.c_size_t
        L.ora_demodulate.argtypes = [C.POINTER(_Demod), C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.ora_demod_init.argtypes = [C.POINTER(_Demod)]
        L.ora_deinterleave_addr.restype = C.c_size_t
        L.ora_deinterleave_addr.argtypes = [C.c_size_t]
        L.ora_viterbi_decode.restype = C.c_int
        L.ora_viterbi_decode.argtypes = [C.c_void_p, C.c_void_p]
        L.ora_frame_decode.restype = C.c_int
        L.ora_frame_decode.argtypes = [C.c_void_p, C.c_void_p]
        L.ora_quantise.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
        L.ora_lfsr_table.argtypes = [C.c_void_p]
        L.ora_run.restype = C.c_int
        L.ora_run.argtypes = [C.POINTER(_Cfg), C.c_void_p, C.c_size_t, C.POINTER(_Result)]
        L.ora_base40_encode.argtypes = [C.c_char_p, C.c_void_p]
        L.ora_bert_frame.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32, C.c_void_p]
        L.ora_encode_frame.argtypes = [C.c_void_p, C.c_void_p]
        L.ora_modulate_frames.restype = C.c_size_t
        L.ora_modulate_frames.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        _lib = L
    return _lib


def have_ref() -> bool:
    return all(os.path.exists(p) for p in (REF_DEMOD, REF_MOD, os.path.join(REF_DIR, "libref_stages.so")))


def ref():
    """The reference's own classes behind oracle/ref_harness.cpp (None when _ref/ is absent)."""
    global _ref
    if _ref is None:
        build()
        if not have_ref():
            return None
        R = C.CDLL(os.path.join(REF_DIR, "libref_stages.so"))
        R.ref_estimate_offset.restype = C.c_double
        R.ref_estimate_offset.argtypes = [C.c_void_p, C.c_size_t]
        R.ref_demodulate_once.restype = C.c_size_t
        R.ref_demodulate_once.argtypes = [C.c_void_p, C.c_size_t, C.c_double, C.c_double, C.c_void_p, C.c_size_t,
                                          C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_size_t)]
        R.ref_run_soft.restype = C.c_size_t
        R.ref_run_soft.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_double, C.c_int, C.c_double,
                                   C.c_void_p, C.c_size_t, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                   C.POINTER(C.c_double), C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        R.ref_run_soft_coherent.restype = C.c_size_t
        R.ref_run_soft_coherent.argtypes = [C.c_void_p, C.c_size_t, C.c_double, C.c_double, C.c_void_p, C.c_size_t,
                                            C.POINTER(C.c_double), C.POINTER(C.c_double)]
        R.ref_track_decode.restype = C.c_size_t
        R.ref_track_decode.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_size_t, C.POINTER(C.c_int)]
        R.ref_frame_decode.restype = C.c_int
        R.ref_frame_decode.argtypes = [C.c_void_p, C.c_void_p]
        R.ref_viterbi_decode.restype = C.c_int
        R.ref_viterbi_decode.argtypes = [C.c_void_p, C.c_void_p]
        R.ref_deinterleave_addr.restype = C.c_size_t
        R.ref_deinterleave_addr.argtypes = [C.c_size_t]
        _ref = R
    return _ref


def _iq(iq: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(iq, dtype=np.int16).reshape(-1)
    assert a.size % 2 == 0
    return a


@dataclass
class RunResult:
    frames: np.ndarray            # [n,134] uint8 (decoded frames, metric >= 0)
    metrics: np.ndarray           # [n] int32
    frame_ready_idx: np.ndarray   # [n] int64 symbol index at which the frame became ready
    soft: np.ndarray              # [n_sym] float64
    events: list                  # [(type, sym_idx, count, corr, raw)]
    chunk_starts: np.ndarray
    n_perfect: int = 0
    n_dropped: int = 0
    est_offset: float = 0.0
    final_freq: float = 0.0
    final_tfreq: float = 0.0
    final_state: int = 0
    extra: dict = field(default_factory=dict)


def run(iq: np.ndarray, streaming: bool, afc_alpha: float = 0.001, init_offset: float | None = None,
        want_soft: bool = True, coherent: bool = False, pll_bw: float = 50.0, coherent_perturb: float = 0.0) -> RunResult:
    """Whole chain through the C restatement (main() drivers, src/opv-demod.cpp:995-1216).
    coherent_perturb: test hook, initial Costas phase offset in radians (0 = the reference's behaviour)."""
    lib().ora_set_codemod_perturb(float(coherent_perturb))
    try:
        return _run(iq, streaming, afc_alpha, init_offset, want_soft, coherent, pll_bw)
    finally:
        lib().ora_set_codemod_perturb(0.0)


def _run(iq, streaming, afc_alpha, init_offset, want_soft, coherent, pll_bw) -> RunResult:
    a = _iq(iq)
    n = a.size // 2
    capf = n // FRAME_SAMPLES + 4
    caps = n // SPS + 16 if want_soft else 0
    frames = np.zeros((capf, FRAME_BYTES), np.uint8)
    metrics = np.zeros(capf, np.int32)
    ready = np.zeros(capf, np.int64)
    soft = np.zeros(max(caps, 1), np.float64)
    cape = 4 * capf + 4096
    ev = (_Event * cape)()
    capc = n // (CHUNK_SAMPLES - 128) + 4
    chunks = np.zeros(capc, np.int64)
    res = _Result()
    res.frames = frames.ctypes.data; res.metrics = metrics.ctypes.data
    res.frame_ready_idx = ready.ctypes.data; res.cap_frames = capf
    res.soft = soft.ctypes.data if want_soft else None; res.cap_soft = caps
    res.events = C.addressof(ev); res.cap_events = cape
    res.chunk_starts = chunks.ctypes.data; res.cap_chunks = capc
    cfg = _Cfg(int(streaming), afc_alpha, int(init_offset is not None), float(init_offset or 0.0),
               int(coherent), float(pll_bw))
    rc = lib().ora_run(C.byref(cfg), a.ctypes.data, n, C.byref(res))
    assert rc == 0
    nf = min(res.n_frames, capf)
    events = [(e.type, e.sym_idx, e.count, e.corr, e.raw) for e in ev[: min(res.n_events, cape)]]
    return RunResult(frames[:nf].copy(), metrics[:nf].copy(), ready[:nf].copy(),
                     soft[: min(res.n_soft, caps)].copy(), events, chunks[: res.n_chunks].copy(),
                     int(res.n_perfect), int(res.n_dropped), res.est_offset, res.final_freq,
                     res.final_tfreq, res.final_state)


def estimate_offset(iq: np.ndarray) -> float:
    a = _iq(iq)
    return lib().ora_estimate_offset(a.ctypes.data, a.size // 2)


def frame_decode(soft2144: np.ndarray):
    s = np.ascontiguousarray(soft2144, np.float64)
    assert s.size == ENCODED_BITS
    out = np.zeros(FRAME_BYTES, np.uint8)
    m = lib().ora_frame_decode(s.ctypes.data, out.ctypes.data)
    return out, m


def viterbi_decode(q2144: np.ndarray):
    q = np.ascontiguousarray(q2144, np.int32)
    bits = np.zeros(FRAME_BITS, np.uint8)
    m = lib().ora_viterbi_decode(q.ctypes.data, bits.ctypes.data)
    return bits, m


def lfsr_table() -> np.ndarray:
    out = np.zeros(FRAME_BYTES, np.uint8)
    lib().ora_lfsr_table(out.ctypes.data)
    return out


def deinterleave_table() -> np.ndarray:
    L = lib()
    return np.array([L.ora_deinterleave_addr(i) for i in range(ENCODED_BITS)], np.int32)


def bert_frames(callsign: str, n: int, token: int = 0xBBAADD, first: int = 0) -> np.ndarray:
    out = np.zeros((n, FRAME_BYTES), np.uint8)
    L = lib()
    for k in range(n):
        L.ora_bert_frame(callsign.encode(), token, first + k, out[k].ctypes.data)
    return out


def modulate(frames: np.ndarray) -> np.ndarray:
    """TX restatement (opv-mod -R semantics): returns int16 [n_samples, 2]."""
    f = np.ascontiguousarray(frames, np.uint8).reshape(-1, FRAME_BYTES)
    n = f.shape[0]
    out = np.zeros((n * FRAME_SAMPLES + 100 * SPS, 2), np.int16)
    ns = lib().ora_modulate_frames(f.ctypes.data, n, out.ctypes.data)
    assert ns == out.shape[0]
    return out


# ---------------------------------------------------------------------------------------------
# the reference binary itself
_EV_RE = [
    (1, re.compile(r"^\[(\d+)\] HUNTING→VERIFYING \(corr=([-\d.]+), raw=([-\d.]+)\)")),
    (2, re.compile(r"^\[(\d+)\] VERIFYING→LOCKED \(frame (\d+)\)")),
    (3, re.compile(r"^\[(\d+)\] LOCKED: sync OK \(corr=([-\d.]+)\)")),
    (4, re.compile(r"^\[(\d+)\] LOCKED: sync MISS #(\d+) \(corr=([-\d.]+)\)")),
    (5, re.compile(r"^\[(\d+)\] LOCKED→HUNTING")),
]


def parse_events(stderr_text: str):
    """Sync events from opv-demod's stderr (src/opv-demod.cpp:651,677,695,699,705) -> [(type, idx, count)]."""
    out = []
    for line in stderr_text.splitlines():
        for typ, rx in _EV_RE:
            m = rx.match(line)
            if m:
                idx = int(m.group(1))
                cnt = int(m.group(2)) if typ in (2, 4) else 0
                out.append((typ, idx, cnt))
                break
    return out


def run_ref_binary(iq: np.ndarray, args=("-r", "-q"), binary: str | None = None):
    """Pipe a capture through the UNMODIFIED reference opv-demod. Returns (frames[n,134], events, exit_code, stderr)."""
    binary = binary or REF_DEMOD
    a = _iq(iq)
    p = subprocess.run([binary, *args], input=a.tobytes(), capture_output=True)
    out = np.frombuffer(p.stdout, np.uint8)
    assert out.size % FRAME_BYTES == 0
    err = p.stderr.decode("utf-8", "replace")
    return out.reshape(-1, FRAME_BYTES).copy(), parse_events(err), p.returncode, err


def run_ref_mod(args, stdin: bytes = b"") -> np.ndarray:
    p = subprocess.run([REF_MOD, *args], input=stdin, capture_output=True, check=True)
    return np.frombuffer(p.stdout, np.int16).reshape(-1, 2).copy()


def ref_run_soft(iq: np.ndarray, streaming: bool, afc_alpha: float = 0.001, init_offset: float | None = None):
    R = ref()
    a = _iq(iq)
    n = a.size // 2
    cap = n // SPS + 16
    soft = np.zeros(cap, np.float64)
    est = C.c_double(); ff = C.c_double(); tf = C.c_double(); nch = C.c_size_t()
    chunks = np.zeros(n // (CHUNK_SAMPLES - 128) + 4, np.int64)
    ns = R.ref_run_soft(a.ctypes.data, n, int(streaming), afc_alpha, int(init_offset is not None),
                        float(init_offset or 0.0), soft.ctypes.data, cap, C.byref(est), C.byref(ff), C.byref(tf),
                        chunks.ctypes.data, chunks.size, C.byref(nch))
    return soft[:ns].copy(), est.value, ff.value, tf.value, chunks[: nch.value].copy()


def ref_run_soft_coherent(iq: np.ndarray, afc_alpha: float = 0.001, pll_bw: float = 50.0):
    """Soft symbols of the reference's CoherentMSKDemodulator (-c, batch)."""
    R = ref()
    a = _iq(iq)
    n = a.size // 2
    cap = n // SPS + 16
    soft = np.zeros(cap, np.float64)
    est = C.c_double(); ff = C.c_double()
    ns = R.ref_run_soft_coherent(a.ctypes.data, n, afc_alpha, pll_bw, soft.ctypes.data, cap, C.byref(est), C.byref(ff))
    return soft[:ns].copy(), est.value, ff.value
