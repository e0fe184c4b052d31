"""CPU test of the PRODUCT's arithmetic: the host/device headers the CUDA kernels are built from
(demod_core.cuh, est_core.cuh, track_core.cuh) compiled for the CPU by tests/hostsim and compared
with the oracle.  This checks the restructured FP64 algorithm (Horner gates, post-sum interpolation,
autocorrelation estimate, event-driven tracker); the GPU tests check the kernels themselves."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


class FR(C.Structure):
    _fields_ = [("start", C.c_int64), ("ready", C.c_int64), ("q", C.c_double)]


class EV(C.Structure):
    _fields_ = [("type", C.c_int32), ("count", C.c_int32), ("idx", C.c_int64), ("corr", C.c_double), ("raw", C.c_double)]


@pytest.fixture(scope="module")
def sim():
    subprocess.run(["make", "-C", os.path.join(HERE, "hostsim")], check=True, stdout=subprocess.DEVNULL)
    H = C.CDLL(os.path.join(HERE, "hostsim", "libhostsim.so"))
    H.hostsim_demod.restype = C.c_size_t
    H.hostsim_demod.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_double, C.c_int, C.c_double, C.c_void_p,
                                C.c_size_t] + [C.POINTER(C.c_double)] * 3
    H.hostsim_demod_warp.restype = C.c_size_t
    H.hostsim_demod_warp.argtypes = H.hostsim_demod.argtypes
    H.hostsim_demod_bank.restype = C.c_size_t
    H.hostsim_demod_bank.argtypes = H.hostsim_demod.argtypes
    H.hostsim_demod_bank_elb.restype = C.c_size_t
    H.hostsim_demod_bank_elb.argtypes = H.hostsim_demod.argtypes
    H.hostsim_demod_coherent.restype = C.c_size_t
    H.hostsim_demod_coherent.argtypes = [C.c_void_p, C.c_size_t, C.c_double, C.c_double, C.c_void_p, C.c_size_t,
                                         C.POINTER(C.c_double), C.POINTER(C.c_double)]
    H.hostsim_track.restype = C.c_size_t
    H.hostsim_track.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                C.POINTER(C.c_size_t), C.POINTER(C.c_int)]
    return H


NAMES = ["clean5", "clean12_call", "awgn14", "awgn8", "awgn4", "cfo_p1200_delay", "cfo_m1900", "random6",
         "dropout_short", "dropout_long", "zeros_gap", "noise_only", "short_lt_chunk", "tiny", "empty"]


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("variant", ["lane", "warp", "bank", "bank_elb"])
def test_device_arithmetic_vs_oracle(name, mode, variant, cases, ora, sim):
    iq = cases[name]
    a = np.ascontiguousarray(iq, np.int16).reshape(-1)
    n = a.size // 2
    ref = ora.run(iq, bool(mode))
    soft = np.zeros(n // 40 + 16)
    est, ff, tf = C.c_double(), C.c_double(), C.c_double()
    fn = {"lane": sim.hostsim_demod, "warp": sim.hostsim_demod_warp, "bank": sim.hostsim_demod_bank, "bank_elb": sim.hostsim_demod_bank_elb}[variant]
    ns = fn(a.ctypes.data, n, mode, 0.001, 0, 0.0, soft.ctypes.data, soft.size, C.byref(est),
                           C.byref(ff), C.byref(tf))
    soft = soft[:ns]
    assert ns == ref.soft.size
    assert est.value == ref.est_offset
    if ns:
        rms = np.sqrt(np.mean(ref.soft ** 2)) + 1e-300
        assert np.max(np.abs(soft - ref.soft)) / rms < 1e-9      # north_star asks 1e-4; FP64 restructuring gives ~1e-12
        assert abs(ff.value - ref.final_freq) < 1e-6
    fr, ev = (FR * 1024)(), (EV * 8192)()
    ne, fs = C.c_size_t(), C.c_int()
    nf = sim.hostsim_track(soft.ctypes.data, ns, fr, 1024, ev, 8192, C.byref(ne), C.byref(fs))
    assert [(e.type, e.idx, e.count) for e in ev[: ne.value]] == [(t, i, c) for (t, i, c, _, _) in ref.events]
    assert fs.value == ref.final_state
    frames = [ora.frame_decode(soft[fr[k].start: fr[k].start + 2144]) for k in range(nf)]
    frames = np.array([f for f, m in frames if m >= 0], np.uint8).reshape(-1, 134)
    assert np.array_equal(frames, ref.frames)                    # bit-exact frames from the device arithmetic


COHERENT_HORIZON = {50.0: 2000, 120.0: 250}  # symbols over which results are pinned, by PLL bandwidth


@pytest.mark.parametrize("name", NAMES)
def test_coherent_device_arithmetic_vs_oracle(name, cases, ora, sim):
    """-c coherent mode (batch only): Horner/sincos restructuring of CoherentMSKDemodulator vs the oracle.

    The reference's Costas loop does not lock on these signals (its own estimator starts it ~1.4 kHz off, the AFC
    runs into the +-2 kHz clamp, and it decodes no error-free frame on any capture, clean ones included), and its
    trajectory is chaotic: a 1-ulp difference grows ~10x per 1,000 symbols (per ~150 with -p 120 -a 0.002).  Results can therefore only be pinned
    over a horizon; beyond it only bit-identical libm sin/cos would reproduce the reference."""
    iq = cases[name]
    a = np.ascontiguousarray(iq, np.int16).reshape(-1)
    n = a.size // 2
    for kw in (dict(), dict(afc_alpha=0.002, pll_bw=120.0)):
        ref = ora.run(iq, False, coherent=True, **kw)
        soft = np.zeros(n // 40 + 16)
        est, ff = C.c_double(), C.c_double()
        ns = sim.hostsim_demod_coherent(a.ctypes.data, n, kw.get("afc_alpha", 0.001), kw.get("pll_bw", 50.0),
                                        soft.ctypes.data, soft.size, C.byref(est), C.byref(ff))
        soft = soft[:ns]
        assert ns == ref.soft.size
        assert est.value == ref.est_offset
        if ns:
            h = min(ns, COHERENT_HORIZON[kw.get("pll_bw", 50.0)])
            rms = np.sqrt(np.mean(ref.soft[:h] ** 2)) + 1e-300
            assert np.max(np.abs(soft[:h] - ref.soft[:h])) / rms < 1e-9
