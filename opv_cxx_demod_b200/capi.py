"""ctypes binding of include/opvd.h (libopvd.so).

This is plumbing only: every call goes straight into the C ABI, whose kernels are hand-written
sm_100a CUDA.  There is deliberately no Python/NumPy/torch implementation of any stage here and no
fallback: if the library is missing or no GPU is usable, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libopvd.so")
CLI_PATH = os.path.join(HERE, "bin", "opv-demod")
BANK_CLI_PATH = os.path.join(os.path.dirname(CLI_PATH), "opv-demod-bank")

FRAME_BYTES = 134
FRAME_SYMBOLS = 2168
ENCODED_BITS = 2144
SPS = 40
CHUNK_SAMPLES = 86720
FRAME_SAMPLES = FRAME_SYMBOLS * SPS
MODE_BATCH, MODE_STREAM = 0, 1
NUM_COUNTERS = 16
COUNTER_NAMES = ["samples", "symbols", "frames_ready", "frames_decoded", "frames_perfect", "frames_dropped",
                 "sync_acq", "sync_ok", "sync_miss", "lost_lock", "bit_errors", "frames_compared", "acs"]

# every symbol include/opvd.h declares (checked by tests/test_abi.py)
EXPORTS = ["opvd_create", "opvd_destroy", "opvd_reset", "opvd_strerror", "opvd_last_cuda_error", "opvd_version",
           "opvd_push_iq", "opvd_push_iq_all", "opvd_attach_device_iq", "opvd_run", "opvd_sync",
           "opvd_poll_frames", "opvd_poll_frames_ready", "opvd_frames_lost", "opvd_poll_events", "opvd_get_soft", "opvd_get_stream_info",
           "opvd_get_counters", "opvd_counters_device_ptr", "opvd_last_run_ms", "opvd_demod_lanes", "opvd_stage_decode",
           "opvd_stage_decode_dev", "opvd_synth_bank", "opvd_bert_check"]


class Config(C.Structure):
    _fields_ = [("n_streams", C.c_int32), ("mode", C.c_int32), ("afc_alpha", C.c_double),
                ("have_init_offset", C.c_int32), ("device", C.c_int32), ("init_offset_hz", C.c_double),
                ("max_samples", C.c_int64), ("max_symbols", C.c_int64), ("max_frames", C.c_int32),
                ("lanes_per_stream", C.c_int32), ("coherent", C.c_int32), ("reserved0", C.c_int32),
                ("pll_bw_hz", C.c_double)]


class Event(C.Structure):
    _fields_ = [("type", C.c_int32), ("count", C.c_int32), ("sym_idx", C.c_int64), ("corr", C.c_double),
                ("raw", C.c_double)]


class FrameInfo(C.Structure):
    _fields_ = [("stream", C.c_int32), ("frame_idx", C.c_int32), ("metric", C.c_int32), ("reserved", C.c_int32),
                ("payload_start", C.c_int64), ("ready_idx", C.c_int64), ("sync_quality", C.c_double)]


class StreamInfo(C.Structure):
    _fields_ = [("est_offset_hz", C.c_double), ("freq_offset_hz", C.c_double), ("timing_freq", C.c_double),
                ("n_symbols", C.c_int64), ("n_samples_used", C.c_int64), ("sync_state", C.c_int32),
                ("frames_ready", C.c_int32), ("done", C.c_int32), ("reserved", C.c_int32)]


class Synth(C.Structure):
    _fields_ = [("n_streams", C.c_int32), ("n_frames", C.c_int32), ("stride_samples", C.c_int64),
                ("n_samples", C.c_int64), ("seed", C.c_uint64), ("scale", C.c_float), ("ebn0_lo_db", C.c_float),
                ("ebn0_hi_db", C.c_float), ("cfo_max_hz", C.c_float), ("frac_delay", C.c_int32),
                ("max_lead", C.c_int32), ("first_stream", C.c_int32), ("reserved", C.c_int32)]


class OpvdError(RuntimeError):
    pass


def build(verbose: bool = False) -> None:
    """Compile libopvd.so and bin/opv-demod in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", os.path.join(HERE, "csrc")], capture_output=not verbose, text=True)
    if r.returncode != 0:
        raise OpvdError("building libopvd.so failed:\n" + (r.stdout or "") + (r.stderr or ""))


_lib = None


def lib() -> C.CDLL:
    """Load libopvd.so.  Fails loudly when it has not been built — there is no other implementation."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OpvdError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(nvcc, sm_100a). opv_cxx_demod_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    H = C.c_void_p
    L.opvd_create.argtypes = [C.POINTER(Config), C.POINTER(H)]
    L.opvd_destroy.argtypes = [H]
    L.opvd_reset.argtypes = [H]
    L.opvd_strerror.restype = C.c_char_p
    L.opvd_strerror.argtypes = [C.c_int]
    L.opvd_last_cuda_error.restype = C.c_char_p
    L.opvd_last_cuda_error.argtypes = [H]
    L.opvd_push_iq.argtypes = [H, C.c_int32, C.c_void_p, C.c_int64]
    L.opvd_push_iq_all.argtypes = [H, C.c_void_p, C.c_int64, C.c_int64]
    L.opvd_attach_device_iq.argtypes = [H, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]
    L.opvd_run.argtypes = [H, C.c_int]
    L.opvd_sync.argtypes = [H]
    L.opvd_poll_frames.argtypes = [H, C.c_int32, C.c_void_p, C.c_void_p]
    L.opvd_poll_frames_ready.argtypes = [H, C.c_int32, C.c_void_p, C.c_void_p]
    L.opvd_frames_lost.argtypes = [H, C.POINTER(C.c_uint64)]
    L.opvd_poll_events.argtypes = [H, C.c_int32, C.c_int32, C.c_void_p]
    L.opvd_get_soft.argtypes = [H, C.c_int32, C.c_int64, C.c_int64, C.c_void_p]
    L.opvd_get_stream_info.argtypes = [H, C.c_int32, C.POINTER(StreamInfo)]
    L.opvd_get_counters.argtypes = [H, C.c_void_p, C.c_int32]
    L.opvd_counters_device_ptr.argtypes = [H, C.POINTER(C.c_void_p)]
    L.opvd_last_run_ms.argtypes = [H, C.c_void_p]
    L.opvd_demod_lanes.argtypes = [H]
    L.opvd_stage_decode.argtypes = [C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    L.opvd_stage_decode_dev.argtypes = [C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    L.opvd_synth_bank.argtypes = [C.c_int32, C.POINTER(Synth), C.c_void_p]
    L.opvd_bert_check.argtypes = [H, C.POINTER(Synth)]
    _lib = L
    return L


def check(rc: int, handle=None, what: str = "") -> int:
    if rc < 0:
        L = lib()
        msg = L.opvd_strerror(rc).decode()
        detail = L.opvd_last_cuda_error(handle).decode() if handle else ""
        raise OpvdError(f"{what}: {msg}" + (f" [{detail}]" if detail else ""))
    return rc
