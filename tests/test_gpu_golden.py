"""GPU: the fused kernels through the C ABI against the REFERENCE'S OWN OUTPUT (golden stdout of
the reference's process.cpp/messageQueue.h run through oracle/_ref, committed under tests/golden),
plus known-answer and edge-case tests."""
import numpy as np
import pytest

import oracle as O
import scanner_b200 as S
from tests import golden_util as GU
from tests import synth

pytestmark = pytest.mark.gpu
G = GU.load()


def format_hits(res, freqs, fs, n, lo, hi):
    """What process.cpp:57 prints for buffers lo..hi-1: (uint64 Hz, power_db) in ascending bin order."""
    out = []
    for b in range(lo, hi):
        c = int(res["hit_count"][b])
        for h in res["hits"][b, :c]:
            out.append((S.hit_frequency(float(freqs[b]), fs, n, int(h["bin"])), float(h["power_db"])))
    return out


@pytest.mark.parametrize("case", [c for c in GU.scan_cases(G) if c["mode"] == 2], ids=lambda c: c["name"])
def test_frequency_mode_matches_reference_stdout(case):
    n = case["n"]
    if n < 256:
        pytest.skip("N below the kernel family")
    window = S.window_build(case["win"], n)
    lo, hi = GU.accepted_range(case)
    with S.SpectrumSense(n, case["fs"], case["enob"], case["thr"], window, sample_kind=case["kind"],
                         correct_dc_offset=case["dc"], max_spectra=hi) as ss:
        res = ss.process(case["raw"])
    got = format_hits(res, case["freqs"], case["fs"], n, lo, hi)
    want = GU.parse_hits(case["text"])
    assert [f for f, _ in got] == [f for f, _ in want]          # same bins, same order, same Hz
    assert max(abs(a - b) for (_, a), (_, b) in zip(got, want)) < 1e-3 + 1e-6


@pytest.mark.parametrize("case", [c for c in GU.scan_cases(G) if c["mode"] == 1], ids=lambda c: c["name"])
def test_time_domain_matches_reference_stdout(case):
    lo, hi = GU.accepted_range(case)
    with S.SpectrumSense(case["n"], case["fs"], case["enob"], case["thr"], None, sample_kind=case["kind"],
                         correct_dc_offset=case["dc"], mode=S.MODE_TIME_DOMAIN, max_spectra=hi) as ss:
        res = ss.process(case["raw"])
    want = GU.parse_time_domain(case["text"])
    trig, mm = res["hit_count"], res["td_max_min"]
    got = [(b - lo, float(mm[b, 0]), float(mm[b, 1])) for b in range(lo, hi) if trig[b]]
    assert [w[0] for w in want] == [g[0] for g in got] and 0 < len(got) < hi - lo
    for w, g in zip(want, got):
        assert abs(w[1] - g[1]) < 1e-5 and abs(w[3] - g[2]) < 1e-5
    otrig, omm = O.time_domain(case["raw"], case["n"], case["enob"], case["kind"], case["dc"], case["thr"])
    np.testing.assert_array_equal(trig, otrig)
    np.testing.assert_allclose(mm, omm, atol=1e-5)


@pytest.mark.parametrize("kind,enob,dc,n", [(1, 8, True, 4096), (2, 12, True, 1000), (3, 12, False, 8192),
                                            (4, 0, False, 2048), (1, 8, False, 65536)])
def test_time_domain_parity_vs_oracle(kind, enob, dc, n):
    raw = synth.make_buffers(kind, n, 24, enob, seed=300 + kind)
    if kind == 2:
        raw[::3, :, :8] = np.iinfo(raw.dtype).max            # some saturated buffers => positive dB
    elif kind != 4:
        raw[::3, :8, :] = np.iinfo(raw.dtype).max
    else:
        raw[::3, :8, :] = 1.5
    _, mm = O.time_domain(raw, n, enob, kind, dc, 0.0)
    thr = float(np.float32(0.5 * np.sort(mm[::3, 0])[0]))
    otrig, omm = O.time_domain(raw, n, enob, kind, dc, thr)
    with S.SpectrumSense(n, 8_000_000, enob, thr, None, sample_kind=kind, correct_dc_offset=dc,
                         mode=S.MODE_TIME_DOMAIN, max_spectra=7) as ss:       # 24 buffers => chunked
        res = ss.process(raw)
    np.testing.assert_array_equal(res["hit_count"], otrig)
    assert 0 < otrig.sum() < 24
    np.testing.assert_allclose(res["td_max_min"], omm, atol=1e-5)


# ---- known-answer tests through the GPU ---------------------------------------------------

@pytest.mark.parametrize("n", [256, 1024, 4096, 16384])
def test_fft_known_answers(n):
    rect = S.window_build(S.WIN_RECTANGULAR, n)
    k, a = n // 5, 0.37
    x = np.zeros((3, n, 2), np.float32)
    x[0, 0, 0] = 1.0                                             # impulse: |X| = 1 everywhere -> 0 dB
    x[1, :, 0] = 1.0                                             # DC: |X[0]| = N
    t = np.arange(n)
    x[2, :, 0] = a * np.cos(2 * np.pi * k * t / n)               # +k exponential lands on bin k (forward sign)
    x[2, :, 1] = a * np.sin(2 * np.pi * k * t / n)
    with S.SpectrumSense(n, 8_000_000, 0, 1e9, rect, sample_kind=S.KIND_FLOAT_COMPLEX, max_spectra=3) as ss:
        db = ss.process(x)["spectra_db"]
    assert np.abs(db[0]).max() < 1e-5
    assert abs(db[1, 0] - 10 * np.log10(n)) < 1e-4 and db[1, 1:].max() < db[1, 0] - 60
    assert np.argmax(db[2]) == k and abs(db[2, k] - 10 * np.log10(n * a)) < 1e-4


def test_zero_input_gives_minus_inf_and_no_hits():
    n = 1024
    raw = np.zeros((2, n, 2), np.int16)
    with S.SpectrumSense(n, 8_000_000, 12, -1000.0, S.window_build(5, n), sample_kind=S.KIND_SHORT_COMPLEX,
                         max_spectra=2) as ss:
        res = ss.process(raw)
    assert np.all(np.isneginf(res["spectra_db"])) and res["hit_count"].sum() == 0     # -inf > thr is false


@pytest.mark.parametrize("n", [256, 2048, 8192])
def test_detection_band_edges_and_dc_hole(n):
    """KAT-4 on the GPU: one strong exponential per spectrum at chosen shifted indices."""
    rect = S.window_build(S.WIN_RECTANGULAR, n)
    use_w, half = S.use_window(0.75, n), n // 2
    idx = [half - use_w - 1, half - use_w, half + use_w, half + use_w + 1] + list(range(half - 5, half + 6))
    t = np.arange(n)
    x = np.zeros((len(idx), n, 2), np.float32)
    for s, i in enumerate(idx):
        j = (i + half) % n
        x[s, :, 0] = np.cos(2 * np.pi * j * t / n)
        x[s, :, 1] = np.sin(2 * np.pi * j * t / n)
    with S.SpectrumSense(n, 8_000_000, 0, 10.0, rect, sample_kind=S.KIND_FLOAT_COMPLEX,
                         max_spectra=len(idx)) as ss:
        res = ss.process(x)
    for s, i in enumerate(idx):
        inside = (half - use_w) <= i <= (half + use_w)
        j = (i + half) % n
        in_hole = j < 4 or (n - j) < 4
        want = 1 if (inside and not in_hole) else 0
        assert res["hit_count"][s] == want, (i, res["hit_count"][s])
        if want:
            assert res["hits"]["bin"][s, 0] == i
            assert res["hit_mask"][s, i >> 5] == np.uint32(1 << (i & 31))
    cand = len(synth.candidate_bins(n, use_w))
    with S.SpectrumSense(n, 8_000_000, 0, -1e9, rect, sample_kind=S.KIND_FLOAT_COMPLEX, max_spectra=1) as ss:
        assert ss.process(x[:1])["hit_count"][0] == cand             # every candidate bin hits


def test_dc_quirk_negative_sum_int8():
    """utility.cpp:49-50: a negative DC sum divides as unsigned -> huge positive dc; the GPU must
    reproduce it bit for bit (compared through the time-domain max and the spectrum)."""
    n = 256
    rng = np.random.default_rng(9)
    raw = rng.integers(-128, -100, (3, n, 2)).astype(np.int8)
    w = S.window_build(5, n)
    use_w = S.use_window(0.75, n)
    truth = O.pipeline(raw, n, 8_000_000, 8, 1, True, 1, 0.0, w, use_w, precision=1, want_f64=True)
    thr = synth.guard_banded_threshold(truth["spectra_db64"], n, use_w, quantile=0.9)
    truth = O.pipeline(raw, n, 8_000_000, 8, 1, True, 1, thr, w, use_w, precision=1, want_f64=True)
    with S.SpectrumSense(n, 8_000_000, 8, thr, w, sample_kind=1, correct_dc_offset=True, max_spectra=3) as ss:
        res = ss.process(raw)
    np.testing.assert_array_equal(res["hit_mask"], truth["hit_mask"])
    strong = truth["spectra_db64"] > truth["spectra_db64"].max() - 40
    assert np.abs(res["spectra_db"] - truth["spectra_db64"])[strong].max() < 1e-3
    assert truth["spectra_db64"].max() > 60          # the quirk's 2^24-sized offset is really there


def test_threshold_update_and_relaunch():
    n = 1024
    raw = synth.make_buffers(3, n, 6, 12, seed=77)
    w = S.window_build(5, n)
    with S.SpectrumSense(n, 8_000_000, 12, 1e9, w, max_spectra=6) as ss:
        assert ss.process(raw)["hit_count"].sum() == 0
        ss.set_threshold(-1e9)
        assert ss.process(raw)["hit_count"].sum() == 6 * len(synth.candidate_bins(n, S.use_window(0.75, n)))
        assert ss.launch_count == 2


def test_invalid_configurations_are_rejected():
    w = S.window_build(5, 1024)
    for kwargs in (dict(sample_count=1000), dict(sample_count=128), dict(enob=9, sample_kind=1),
                   dict(enob=0, sample_kind=3), dict(sample_kind=7), dict(max_spectra=0)):
        args = dict(sample_count=1024, sample_rate=8_000_000, enob=12, threshold=1.0, window=w, sample_kind=3)
        args.update(kwargs)
        if args["sample_count"] != 1024:
            args["window"] = np.ones(args["sample_count"], np.float32)
        with pytest.raises(S.ScannerError) as e:
            S.SpectrumSense(**args)
        assert e.value.status == 1
    with S.SpectrumSense(1024, 8_000_000, 12, 1.0, w, max_spectra=4, ticket_slots=1) as ss:
        raw = np.zeros((4, 1024, 2), np.int16)
        t = ss.submit(raw.ctypes.data, 4)
        with pytest.raises(S.ScannerError) as e:
            ss.submit(raw.ctypes.data, 4)               # slot still in flight
        assert e.value.status == 5
        ss.collect(t)
        with pytest.raises(S.ScannerError) as e:
            ss.submit(raw.ctypes.data, 5)               # over capacity
        assert e.value.status == 4
