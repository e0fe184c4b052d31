"""CPU tests: the oracle (oracle/opv_oracle.c) against the reference's golden outputs and, when the
reference build is present (authoring container / shipped oracle/_ref), against the reference itself."""
import hashlib
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "reference_golden.json")))["cases"]
STAGE = np.load(os.path.join(HERE, "golden", "stage_vectors.npz"))


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("key", sorted(GOLD))
def test_oracle_matches_reference_golden(key, cases, ora):
    name, mode = key.split("/")
    iq = cases[name]
    g = GOLD[key]
    assert _sha(iq) == g["capture_sha256"], "capture generator drifted from the one the golden was made with"
    r = ora.run(iq, mode == "stream", coherent=(mode == "coherent"))
    assert r.frames.shape[0] == g["n_frames"]
    assert _sha(r.frames) == g["frames_sha256"]                      # stdout bytes of the reference binary
    assert [[t, i, c] for (t, i, c, _, _) in r.events] == g["events"]  # its stderr transitions
    assert (0 if r.frames.shape[0] > 0 else 1) == g["exit_code"]
    assert r.soft.size == g["n_soft"]
    assert _sha(r.soft) == g["soft_sha256"]                          # soft symbols bit-identical
    assert r.est_offset == g["est_offset"]
    assert r.final_freq == g["final_freq"]
    assert [int(c) for c in r.chunk_starts] == g["chunk_starts"]
    if g["summary"]:
        assert g["summary"] == f"Summary: {len(r.frames)} frames ({r.n_perfect} perfect, {len(r.frames) - r.n_perfect} errors)"


def test_stage_vectors(ora):
    for p, f, m in zip(STAGE["payloads"], STAGE["frames"], STAGE["metrics"]):
        got, metric = ora.frame_decode(p)
        assert metric == m
        if m >= 0:
            assert np.array_equal(got, f)
    for q, b, m in zip(STAGE["vit_in"], STAGE["vit_bits"], STAGE["vit_metric"]):
        bits, metric = ora.viterbi_decode(q)
        assert metric == m and np.array_equal(bits, b)


def test_known_answer_constants(ora):
    # constants derivable from the reference code
    lf = ora.lfsr_table()
    assert lf[:16].tolist() == [0xFF, 0x1A, 0xAF, 0x66, 0x52, 0x23, 0x1E, 0x10, 0xA0, 0xF9, 0xFA, 0x8A, 0x98, 0x67, 0x7D, 0xD2]
    assert lf[-1] == 0x31
    d = ora.deinterleave_table()
    assert d[:12].tolist() == [7, 68, 129, 206, 267, 328, 405, 466, 543, 604, 665, 742]
    assert d[32:36].tolist() == [6, 67, 128, 205]
    assert sorted(d.tolist()) == list(range(2144))
    f = ora.bert_frames("W5NYV", 1)[0]
    assert f[:12].tolist() == [0x00, 0x00, 0x03, 0x74, 0x26, 0x97, 0xBB, 0xAA, 0xDD, 0, 0, 0]
    assert f[12:].tolist() == [i & 0xFF for i in range(122)]


def test_clean_loopback_decodes_bert_payloads(cases, ora):
    # Makefile:23-33 style loopback: every frame of a clean capture equals its BERT payload
    for streaming in (False, True):
        r = ora.run(cases["clean12_call"], streaming)
        assert np.array_equal(r.frames, ora.bert_frames("KB5MU", 12, first=250))
        assert r.n_perfect == 12
        assert r.events[0][:2] == (1, 23) and r.events[1][:2] == (2, 2167)


@pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(HERE), "oracle", "_ref", "opv-demod")),
                    reason="reference build (oracle/_ref) not present")
@pytest.mark.parametrize("name", ["clean5", "awgn4", "cfo_m1900", "dropout_long", "zeros_gap", "short_lt_chunk"])
def test_oracle_vs_reference_live(name, cases, ora):
    iq = cases[name]
    for streaming in (False, True):
        r = ora.run(iq, streaming)
        fb, evb, rc, _ = ora.run_ref_binary(iq, ["-r"] + (["-s"] if streaming else []))
        assert np.array_equal(fb, r.frames)
        assert [(t, i, c) for (t, i, c, _, _) in r.events] == evb
        soft_ref, est, ff, tf, chunks = ora.ref_run_soft(iq, streaming)
        assert np.array_equal(soft_ref, r.soft) and est == r.est_offset and ff == r.final_freq


@pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(HERE), "oracle", "_ref", "opv-demod")),
                    reason="reference build (oracle/_ref) not present")
@pytest.mark.parametrize("name", ["clean5", "awgn14", "awgn8", "cfo_p1200_delay", "dropout_long", "zeros_gap", "tiny"])
def test_coherent_oracle_vs_reference_live(name, cases, ora):
    """-c (CoherentMSKDemodulator, batch only): restatement bit-identical to the reference class and binary."""
    iq = cases[name]
    for kw in (dict(), dict(pll_bw=120.0, afc_alpha=0.002)):
        r = ora.run(iq, False, coherent=True, **kw)
        args = ["-c", "-r"] + (["-p", str(kw["pll_bw"]), "-a", str(kw["afc_alpha"])] if kw else [])
        fb, evb, rc, _ = ora.run_ref_binary(iq, args)
        assert np.array_equal(fb, r.frames)
        assert [(t, i, c) for (t, i, c, _, _) in r.events] == evb
        soft_ref, est, ff = ora.ref_run_soft_coherent(iq, kw.get("afc_alpha", 0.001), kw.get("pll_bw", 50.0))
        assert np.array_equal(soft_ref, r.soft) and est == r.est_offset and ff == r.final_freq


@pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(HERE), "oracle", "_ref", "opv-mod")),
                    reason="reference build (oracle/_ref) not present")
def test_tx_restatement_vs_reference_mod(ora):
    from tools import captures as cap

    assert np.array_equal(ora.run_ref_mod(["-S", "W5NYV", "-B", "3"]), cap.clean_bert(3))
    iq, frames = cap.clean_random(2, 5)
    assert np.array_equal(ora.run_ref_mod(["-R"], stdin=frames.tobytes()), iq)


def test_coherent_mode_is_chaotic_in_the_last_bit(cases, ora):
    """`-c` (CoherentMSKDemodulator): why its frames cannot be pinned on long captures.  The reference's Costas loop
    never locks on these signals, and its trajectory is chaotic: offsetting the initial loop phase of the ORACLE ITSELF
    by 2e-16 rad (one ulp of a phase near 1 rad — the size of a libm-vs-libdevice sin/cos difference) leaves the soft
    symbols identical for a few thousand symbols and then turns them into a different sequence; every capture on which
    the coherent path emits frames at all emits DIFFERENT frames.  Bit-identical frames from `-c` therefore need a
    bit-identical libm, not just a correct algorithm; the GPU tests pin `-c` over the stable horizon and compare frames
    on captures shorter than it (tests/test_gpu_parity.py::test_coherent_mode_vs_oracle)."""
    horizons, changed, emitting = [], 0, 0
    for name, iq in cases.items():
        if iq.shape[0] < 40 * 3000:
            continue
        r0 = ora.run(iq, False, coherent=True)
        r1 = ora.run(iq, False, coherent=True, coherent_perturb=2e-16)
        rms = np.sqrt(np.mean(r0.soft ** 2)) + 1e-300
        bad = np.nonzero(np.abs(r1.soft - r0.soft) / rms > 1e-9)[0]
        if bad.size:
            horizons.append(int(bad[0]))
        if r0.frames.shape[0]:
            emitting += 1
            changed += int(not np.array_equal(r0.frames, r1.frames))
    assert len(horizons) >= 8 and 1500 <= min(horizons) and max(horizons) <= 8000, horizons
    assert emitting >= 5 and changed >= emitting - 1, (emitting, changed)   # zeros_gap survives this particular offset
    # and the hook itself is inert at zero
    iq = cases["awgn8"]
    assert np.array_equal(ora.run(iq, False, coherent=True).soft, ora.run(iq, False, coherent=True, coherent_perturb=0.0).soft)


def test_oracle_is_reentrant(ora):
    """The parity tests run the oracle over hundreds of captures on a thread pool (the C code runs outside the GIL):
    results must not depend on what the other threads are doing (no static scratch buffers).  Random payloads: with
    the BERT captures every thread decodes the same frames and a shared buffer goes unnoticed."""
    from concurrent.futures import ThreadPoolExecutor
    from tools import captures as cap

    caps = [cap.clean_random(6, seed)[0] for seed in range(8)]
    serial = [ora.run(c, True) for c in caps]
    assert all(r.frames.shape[0] == 6 for r in serial)
    jobs = list(range(8)) * 12
    with ThreadPoolExecutor(max_workers=16) as ex:
        got = list(ex.map(lambda k: ora.run(caps[k], True), jobs))
    for k, r in zip(jobs, got):
        assert np.array_equal(r.frames, serial[k].frames) and np.array_equal(r.soft, serial[k].soft), k
        assert r.events == serial[k].events
