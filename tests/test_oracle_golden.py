"""CPU: the oracle restatement against golden vectors produced by the reference's own sources
(oracle/_ref, see tests/golden/make_golden.py), plus the known-answer tests of SURVEY.md 8c."""
import numpy as np
import pytest

import oracle as O
from tests import golden_util as GU
from tests import synth

G = GU.load()


@pytest.mark.parametrize("row", [tuple(r) for r in G["conv_cases"]])
def test_converters_bit_exact(row):
    name, kind, n, enob, dc = row[0], int(row[1]), int(row[2]), int(row[3]), bool(int(row[4]))
    raw, want = G[f"conv_{name}_raw"], G[f"conv_{name}_out"]
    for b in range(raw.shape[0]):
        got = O.convert(kind, raw[b], n, enob, dc)
        np.testing.assert_array_equal(got.view(np.uint32), want[b].view(np.uint32))


def test_converter_probe_values():
    """SURVEY.md KAT-1 [probe] values, as the reference itself computes them."""
    assert G["conv_i8_probe_out"].ravel().tolist() == [-0.078125, -0.15625]        # enob 8 => scale -1/128
    assert G["conv_i16_probe_out"].ravel().tolist() == [0.048828125, 0.09765625]
    assert G["conv_i16_negsum_quirk_out"][0, 0].tolist() == [-524288.0, -524288.0]  # unsigned-division quirk
    assert np.all(G["conv_i8_dc_mean_out"][0, 0] == [0.0078125, 0.0078125])


def test_magnitude_db_both_header_readings_bit_exact():
    x = G["mag_in"]
    for variant, key in ((0, "mag_out_math"), (1, "mag_out_cmath")):
        got, want = O.magnitude_db(x, variant), G[key]
        np.testing.assert_array_equal(got.view(np.uint32), want.view(np.uint32))
    a, b = G["mag_out_math"][1:], G["mag_out_cmath"][1:]
    d = np.abs(a.astype(np.float64) - b)
    assert np.nanmax(d / np.spacing(np.abs(a))) <= 2          # the two readings differ by <= 2 ulp (SURVEY defect 3)
    assert G["mag_out_math"][0] == -np.inf and G["mag_out_math"][1] == 0.0
    assert abs(float(G["mag_out_math"][2]) - 6.98969984) < 1e-6     # KAT-2: (3,4)


def test_frequency_tables():
    for i, c in enumerate(G["ft_cases"]):
        got = O.frequency_table(int(c[0]), c[1], c[2], c[3], c[4])
        np.testing.assert_array_equal(got, G[f"ft_{i}"])
    assert len(G["ft_0"]) == 50 and G["ft_0"][0] == 2407500000 and G["ft_0"][1] - G["ft_0"][0] == 15000000
    assert len(G["ft_1"]) == 24 and G["ft_1"][0] == 1021000000
    assert len(G["ft_2"]) == 133


@pytest.mark.parametrize("case", [c for c in GU.scan_cases(G) if c["mode"] == 2], ids=lambda c: c["name"])
def test_pipeline_matches_reference_output(case):
    """Reference stdout ('freq %lu power_db %f', process.cpp:57) for the accepted buffers vs the oracle:
    same frequencies in the same order; power within 1e-3 dB (+ the 6-decimal print)."""
    n, kind = case["n"], case["kind"]
    window = O.window_build(case["win"], n)
    use_w = O.use_window(0.75, n)
    lo, hi = GU.accepted_range(case)
    want = GU.parse_hits(case["text"])
    for precision in (0, 1):
        res = O.pipeline(case["raw"], n, case["fs"], case["enob"], kind, case["dc"], 1, case["thr"], window,
                         use_w, precision=precision)
        got = []
        for b in range(lo, hi):
            cnt, _, bins = O.detect(res["spectra_db"][b], use_w, 4, case["thr"])
            for i in bins:
                got.append((O.hit_frequency(case["freqs"][b], case["fs"], n, int(i)),
                            float(res["spectra_db"][b][(int(i) + n // 2) % n])))
        assert [f for f, _ in got] == [f for f, _ in want]
        assert len(want) > 0
        assert max(abs(a - b) for (_, a), (_, b) in zip(got, want)) < 1e-3 + 1e-6
    assert case["text"].count("Start scan at") == case["sweeps"] - 1     # first sweep is dropped


@pytest.mark.parametrize("case", [c for c in GU.scan_cases(G) if c["mode"] == 1], ids=lambda c: c["name"])
def test_time_domain_matches_reference_output(case):
    lo, hi = GU.accepted_range(case)
    trig, mm = O.time_domain(case["raw"], case["n"], case["enob"], case["kind"], case["dc"], case["thr"])
    want = GU.parse_time_domain(case["text"])
    got = [(b - lo, float(mm[b, 0]), float(case["freqs"][b]), float(mm[b, 1])) for b in range(lo, hi) if trig[b]]
    assert [w[0] for w in want] == [g[0] for g in got] and 0 < len(want) < hi - lo
    for w, g in zip(want, got):
        assert abs(w[1] - g[1]) < 2e-6 and w[2] == g[2] and abs(w[3] - g[3]) < 2e-6


# ---- known-answer tests -----------------------------------------------------------------

@pytest.mark.parametrize("n", [256, 1024, 2048, 8192, 65536])
def test_fft_kat(n):
    rng = np.random.default_rng(n)
    imp = np.zeros(n, np.complex64); imp[0] = 1
    np.testing.assert_allclose(O.fft_f32(imp), np.ones(n), atol=0)
    dc = np.ones(n, np.complex64)
    want = np.zeros(n); want[0] = n
    np.testing.assert_allclose(np.abs(O.fft_f32(dc)), want, atol=1e-3)
    k, a = n // 5, 0.37
    tone = (a * np.exp(2j * np.pi * k * np.arange(n) / n)).astype(np.complex64)
    X = O.fft_f64(tone)
    assert abs(abs(X[k]) - n * a) < 1e-4 * n * a and np.argmax(np.abs(X)) == k     # forward sign: +k lands on bin k
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    ref = np.fft.fft(x.astype(np.complex128))
    assert np.abs(O.fft_f64(x) - ref).max() < 1e-9 * np.abs(ref).max()
    assert np.abs(O.fft_f32(x) - ref).max() < 1e-5 * np.abs(ref).max()
    assert abs(np.sum(np.abs(O.fft_f64(x)) ** 2) / n - np.sum(np.abs(x.astype(np.complex128)) ** 2)) < 1e-6 * n   # Parseval


@pytest.mark.parametrize("n,count", [(1024, 762), (2048, 1530), (4096, 3066), (8192, 6138)])
def test_detect_indexing(n, count):
    """KAT-4: inclusive band edges, the 7-bin DC hole, candidate counts."""
    use_w = O.use_window(0.75, n)
    cnt, mask, bins = O.detect(np.full(n, 1.0, np.float32), use_w, 4, 0.0)
    assert cnt == count == len(synth.candidate_bins(n, use_w))
    half = n // 2
    assert bins[0] == half - use_w and bins[-1] == half + use_w
    hole = set(range(half - 3, half + 4))        # j in {N-3,N-2,N-1,0,1,2,3}
    assert hole.isdisjoint(set(bins.tolist())) and (half - 4) in bins and (half + 4) in bins
    db = np.full(n, -1.0, np.float32); db[5] = 1.0      # strict '>' and fftshift: FFT bin j=5 <-> i = half+5
    cnt, mask, bins = O.detect(db, use_w, 4, 0.0)
    assert cnt == 1 and bins[0] == half + 5 and mask[(half + 5) >> 5] == 1 << ((half + 5) & 31)
    assert O.detect(np.full(n, 0.0, np.float32), use_w, 4, 0.0)[0] == 0
    assert O.detect(np.full(n, np.nan, np.float32), use_w, 4, 0.0)[0] == 0


def test_hz_mapping():
    """KAT-5: truncating uint32 fs/N and fs/2 (process.cpp:38-39,55)."""
    assert 20_000_000 // 2048 == 9765
    assert O.hit_frequency(2.4075e9, 20_000_000, 2048, 0) == 2397500000
    assert O.hit_frequency(2.4075e9, 20_000_000, 2048, 100) == 2397500000 + 100 * 9765
    assert O.hit_frequency(303e6, 8_000_000, 1024, 512) == 299000000 + 512 * 7812


def test_k1_averaging_is_identity():
    """KAT-7: K = 1 averaging is bitwise the un-averaged path."""
    n = 1024
    raw = synth.make_buffers(3, n, 4, 12, seed=5)
    w = O.window_build(5, n)
    a = O.pipeline(raw, n, 8_000_000, 12, 3, False, 1, 10.0, w, O.use_window(0.75, n), precision=0)
    for b in range(4):
        iq = O.window_apply(O.convert(3, raw[b], n, 12, False), w)
        db = O.magnitude_db(O.fft_f32(iq[:, 0] + 1j * iq[:, 1]))
        np.testing.assert_array_equal(a["spectra_db"][b].view(np.uint32), db.view(np.uint32))


def test_batched_fft_of_the_timed_baseline_is_bit_identical_to_the_scalar_plan():
    """bench.py's CPU arm runs its FFTs eight at a time (one per SIMD lane); same roundings, so same bits."""
    rng = np.random.default_rng(11)
    for n in (256, 1024, 2048, 8192):
        x = rng.standard_normal((8, n, 2)).astype(np.float32)
        got = O.fft_f32_batch8(x)
        for l in range(8):
            want = O.fft_f32(x[l].view(np.complex64).reshape(n)).view(np.float32).reshape(n, 2)
            assert np.array_equal(got[l].view(np.uint32), want.view(np.uint32))


def test_batched_baseline_counts_the_same_hits():
    from tests import synth
    n = 2048
    raw = synth.make_buffers(1, n, 37, 8, seed=77)          # 4 groups of 8 + 5 leftovers
    w, use_w = O.window_build(5, n), O.use_window(0.75, n)
    a = O.bench(raw, n, 20_000_000, 8, 1, True, 1, 8.0, w, use_w, repeats=1, threads=3, faithful=True)
    b = O.bench(raw, n, 20_000_000, 8, 1, True, 1, 8.0, w, use_w, repeats=1, threads=3, faithful=2)
    res = O.pipeline(raw, n, 20_000_000, 8, 1, True, 1, 8.0, w, use_w, precision=0)
    assert a[1] == b[1] == int(res["hit_count"].sum()) > 0
