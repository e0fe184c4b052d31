"""Host-side mirror of the reference's opv-demod objects over the libopvd C ABI.

`DemodBank` plays the role of the (MSKDemodulatorAFC, SyncTracker, FrameDecoder) triple that the
reference's main() builds per process (/root/reference/src/opv-demod.cpp:999-1001, :1164-1183) for
`n_streams` independent streams on one GPU.  Method names follow the reference's verbs; all
arithmetic happens in the CUDA kernels behind the ABI.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import capi
from .capi import (CHUNK_SAMPLES, COUNTER_NAMES, ENCODED_BITS, FRAME_BYTES, FRAME_SAMPLES, MODE_BATCH,
                   MODE_STREAM, NUM_COUNTERS, OpvdError)

SYNC_STATE_NAMES = {0: "HUNTING", 1: "VERIFYING", 2: "LOCKED"}


@dataclass
class Frames:
    data: np.ndarray           # [n,134] uint8
    stream: np.ndarray         # [n] int32
    frame_idx: np.ndarray      # [n] int32
    metric: np.ndarray         # [n] int32
    payload_start: np.ndarray  # [n] int64
    ready_idx: np.ndarray      # [n] int64
    sync_quality: np.ndarray   # [n] float64

    def of_stream(self, s: int) -> np.ndarray:
        return self.data[self.stream == s]


class DemodBank:
    """n_streams x (demodulator + sync tracker + frame decoder) on one GPU."""

    def __init__(self, n_streams: int, streaming: bool = False, afc_alpha: float = 0.001,
                 init_offset_hz: float | None = None, device: int = -1, max_samples: int = 0,
                 max_symbols: int = 0, max_frames: int = 0, lanes_per_stream: int = 0, coherent: bool = False,
                 pll_bw_hz: float = 50.0):
        self._lib = capi.lib()
        self.n_streams = int(n_streams)
        self.streaming = bool(streaming)
        cfg = capi.Config(self.n_streams, MODE_STREAM if streaming else MODE_BATCH, float(afc_alpha),
                          int(init_offset_hz is not None), int(device), float(init_offset_hz or 0.0),
                          int(max_samples), int(max_symbols), int(max_frames), int(lanes_per_stream),
                          int(bool(coherent)), 0, float(pll_bw_hz))
        self._h = C.c_void_p()
        capi.check(self._lib.opvd_create(C.byref(cfg), C.byref(self._h)), None, "opvd_create")
        self._keepalive = None

    # -- lifetime ------------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.opvd_destroy(self._h)
            self._h = None
        self._keepalive = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _ck(self, rc: int, what: str) -> int:
        return capi.check(rc, self._h, what)

    # -- input ---------------------------------------------------------------------------------
    @staticmethod
    def _as_iq(iq) -> np.ndarray:
        a = np.asarray(iq)
        if a.dtype != np.int16:
            raise TypeError("I/Q samples must be int16 (interleaved I,Q)")
        return np.ascontiguousarray(a)

    def push_iq(self, stream: int, iq) -> None:
        """Append samples to one stream (host memory): the reference's stdin read loop."""
        a = self._as_iq(iq).reshape(-1)
        self._ck(self._lib.opvd_push_iq(self._h, int(stream), a.ctypes.data, a.size // 2), "opvd_push_iq")

    def push_iq_all(self, iq) -> None:
        """Append [n_streams, n_samples, 2] int16 to every stream with one strided H2D copy."""
        a = self._as_iq(iq)
        if a.ndim != 3 or a.shape[0] != self.n_streams or a.shape[2] != 2:
            raise ValueError("expected int16 [n_streams, n_samples, 2]")
        self._ck(self._lib.opvd_push_iq_all(self._h, a.ctypes.data, a.shape[1], a.shape[1]), "opvd_push_iq_all")

    def push_iq_host_ptr(self, ptr: int, n_samples: int, host_stride_samples: int) -> None:
        """Same as push_iq_all for a raw (e.g. pinned) host pointer laid out [n_streams][host_stride]."""
        self._ck(self._lib.opvd_push_iq_all(self._h, C.c_void_p(ptr), int(n_samples), int(host_stride_samples)),
                 "opvd_push_iq_all")

    def attach_device_iq(self, dev_ptr: int, stride_samples: int, n_samples, keepalive=None) -> None:
        """Use captures resident in HBM: [n_streams][stride_samples] packed int16 I/Q words."""
        if np.isscalar(n_samples):
            self._ck(self._lib.opvd_attach_device_iq(self._h, C.c_void_p(dev_ptr), int(stride_samples), None,
                                                     int(n_samples)), "opvd_attach_device_iq")
        else:
            n = np.ascontiguousarray(n_samples, np.int64)
            assert n.size == self.n_streams
            self._ck(self._lib.opvd_attach_device_iq(self._h, C.c_void_p(dev_ptr), int(stride_samples),
                                                     n.ctypes.data, 0), "opvd_attach_device_iq")
        self._keepalive = keepalive

    # -- processing ----------------------------------------------------------------------------
    def run(self, final: bool = True, sync: bool = True) -> None:
        self._ck(self._lib.opvd_run(self._h, int(final)), "opvd_run")
        if sync:
            self.sync()

    def reset(self) -> None:
        """Fresh demodulator / tracker / decoder state for every stream; buffers and attached captures are kept."""
        self._ck(self._lib.opvd_reset(self._h), "opvd_reset")

    def sync(self) -> None:
        self._ck(self._lib.opvd_sync(self._h), "opvd_sync")

    def last_run_ms(self) -> dict:
        ms = (C.c_float * 5)()
        self._ck(self._lib.opvd_last_run_ms(self._h, ms), "opvd_last_run_ms")
        return dict(zip(("estimate", "demod", "track", "decode", "total"), [float(x) for x in ms]))

    def demod_variant(self) -> str:
        """Name of the demodulator kernel this bank runs (automatic selection resolved)."""
        lanes = int(self._lib.opvd_demod_lanes(self._h))
        return {32: "demod_warp_kernel", 96: "demod_bank_kernel"}.get(
            lanes, f"demod(lanes={lanes})")

    # -- output --------------------------------------------------------------------------------
    _INFO_DTYPE = np.dtype([("stream", np.int32), ("frame_idx", np.int32), ("metric", np.int32), ("reserved", np.int32),
                            ("payload_start", np.int64), ("ready_idx", np.int64), ("sync_quality", np.float64)])

    def poll_frames(self, max_frames: int | None = None, wait: bool = True) -> Frames:
        """Frames decoded since the last poll: the reference's cout.write stream.  wait=True waits for the enqueued runs;
        wait=False returns what the runs that have already finished decoded (opvd_poll_frames_ready)."""
        fn = self._lib.opvd_poll_frames if wait else self._lib.opvd_poll_frames_ready
        assert self._INFO_DTYPE.itemsize == C.sizeof(capi.FrameInfo)
        chunks, infos = [], []
        cap = 65536
        remaining = max_frames
        while True:
            want = cap if remaining is None else min(cap, remaining)
            if want <= 0:
                break
            buf = np.empty((want, FRAME_BYTES), np.uint8)
            info = np.empty(want, self._INFO_DTYPE)
            n = self._ck(fn(self._h, want, buf.ctypes.data, info.ctypes.data), "opvd_poll_frames")
            if n:
                chunks.append(buf[:n])
                infos.append(info[:n])
            if remaining is not None:
                remaining -= n
            if n < want:
                break
        data = np.concatenate(chunks) if chunks else np.zeros((0, FRAME_BYTES), np.uint8)
        inf = np.concatenate(infos) if infos else np.zeros(0, self._INFO_DTYPE)
        return Frames(data, inf["stream"].copy(), inf["frame_idx"].copy(), inf["metric"].copy(), inf["payload_start"].copy(),
                      inf["ready_idx"].copy(), inf["sync_quality"].copy())

    def frames_lost(self) -> int:
        """Frames overwritten in the device log before they were polled (0 unless polls are too rare)."""
        v = C.c_uint64()
        self._ck(self._lib.opvd_frames_lost(self._h, C.byref(v)), "opvd_frames_lost")
        return int(v.value)

    def poll_events(self, stream: int):
        """[(type, sym_idx, count, corr, raw)] — the tracker's stderr lines of the reference."""
        out = []
        cap = 1024
        ev = (capi.Event * cap)()
        while True:
            n = self._ck(self._lib.opvd_poll_events(self._h, int(stream), cap, ev), "opvd_poll_events")
            out.extend((e.type, e.sym_idx, e.count, e.corr, e.raw) for e in ev[:n])
            if n < cap:
                break
        return out

    def get_soft(self, stream: int, first_sym: int = 0, n: int | None = None) -> np.ndarray:
        if n is None:
            n = self.stream_info(stream)["n_symbols"] - first_sym
        out = np.zeros(max(int(n), 0), np.float64)
        m = self._ck(self._lib.opvd_get_soft(self._h, int(stream), int(first_sym), out.size, out.ctypes.data),
                     "opvd_get_soft")
        return out[:m]

    def stream_info(self, stream: int) -> dict:
        si = capi.StreamInfo()
        self._ck(self._lib.opvd_get_stream_info(self._h, int(stream), C.byref(si)), "opvd_get_stream_info")
        d = {k: getattr(si, k) for k, _ in si._fields_ if k != "reserved"}
        d["sync_state_name"] = SYNC_STATE_NAMES.get(si.sync_state, "?")
        return d

    def counters(self) -> dict:
        a = np.zeros(NUM_COUNTERS, np.uint64)
        self._ck(self._lib.opvd_get_counters(self._h, a.ctypes.data, NUM_COUNTERS), "opvd_get_counters")
        return {k: int(a[i]) for i, k in enumerate(COUNTER_NAMES)}

    def counters_device_ptr(self) -> int:
        p = C.c_void_p()
        self._ck(self._lib.opvd_counters_device_ptr(self._h, C.byref(p)), "opvd_counters_device_ptr")
        return int(p.value)

    def bert_check(self, synth: "capi.Synth") -> None:
        self._ck(self._lib.opvd_bert_check(self._h, C.byref(synth)), "opvd_bert_check")


def stage_decode(payloads, device: int = -1):
    """FrameDecoder::decode seam (src/opv-demod.cpp:854-898): [n,2144] float64 -> ([n,134] uint8, [n] metric)."""
    p = np.ascontiguousarray(payloads, np.float64).reshape(-1, ENCODED_BITS)
    n = p.shape[0]
    frames = np.zeros((n, FRAME_BYTES), np.uint8)
    metrics = np.zeros(n, np.int32)
    capi.check(capi.lib().opvd_stage_decode(int(device), p.ctypes.data, n, frames.ctypes.data, metrics.ctypes.data),
               None, "opvd_stage_decode")
    return frames, metrics


def make_synth(n_streams: int, n_frames: int, stride_samples: int, n_samples: int, seed: int = 1,
               scale: float = 0.25, ebn0_lo_db: float = -1000.0, ebn0_hi_db: float = -1000.0,
               cfo_max_hz: float = 0.0, frac_delay: bool = False, max_lead: int = 0,
               first_stream: int = 0) -> "capi.Synth":
    return capi.Synth(int(n_streams), int(n_frames), int(stride_samples), int(n_samples), int(seed), float(scale),
                      float(ebn0_lo_db), float(ebn0_hi_db), float(cfo_max_hz), int(bool(frac_delay)), int(max_lead),
                      int(first_stream), 0)


def synth_bank(dev_ptr: int, synth: "capi.Synth", device: int = -1) -> None:
    """Fill a device buffer [n_streams][stride] with a synthetic OPV channel bank (measurement aid)."""
    capi.check(capi.lib().opvd_synth_bank(int(device), C.byref(synth), C.c_void_p(dev_ptr)), None, "opvd_synth_bank")


__all__ = ["DemodBank", "Frames", "stage_decode", "make_synth", "synth_bank", "OpvdError", "CHUNK_SAMPLES",
           "FRAME_SAMPLES", "FRAME_BYTES", "ENCODED_BITS"]
