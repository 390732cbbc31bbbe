"""HackRF sweep path: frame-header pre-pass + rx-callback bookkeeping.

Golden: tests/golden/hackrf_vectors.npz = outputs of the reference's own hackRFSource.cpp compiled
unmodified against a libhackrf shim (tests/golden/make_golden_hackrf.py).  CPU tests pin the oracle
restatement and the host class against it; GPU tests pin the device pre-pass kernel (bit exact) and the
whole replay (scan_b200 hackrf) against the reference's stdout."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle as O                      # noqa: E402
from tests import golden_util as GU     # noqa: E402
from tests import hackrf_stream as HS   # noqa: E402

G = np.load(os.path.join(ROOT, "tests", "golden", "hackrf_vectors.npz"), allow_pickle=False)
TOOL = os.path.join(ROOT, "scanner_b200", "scan_b200")
SELFTEST = os.path.join(ROOT, "scanner_b200", "host_selftest")
ENV = dict(os.environ, TZ="UTC")
CASES = [str(c) for c in G["prepass_cases"]]
N, FS, START, STOP, THR, ITERATIONS, VALID, TPS = (lambda p: (int(p[0]), int(p[1]), float(p[2]), float(p[3]),
                                                             float(p[4]), int(p[5]), int(p[6]), int(p[7])))(G["sweep_params"])
_, OFFSET = HS.scan_parameters(FS, START)


def centre(freq_hz):
    """double(frequencyHz + m_scanOffset), uint64 arithmetic (hackRFSource.cpp:221)."""
    return np.array([float((int(f) + OFFSET) & 0xFFFFFFFFFFFFFFFF) for f in freq_hz])


def mismatch_lines(text):
    return [l for l in str(text).splitlines() if l.startswith("interpolateSamples")]


def test_generator_reproduces_the_committed_inputs():
    for name, valid, transfers in HS.prepass_cases():
        assert np.array_equal(transfers, G[f"prepass_{name}_in"]), name


@pytest.mark.parametrize("name", CASES)
def test_oracle_prepass_matches_reference(name):
    tin, want, wfreq = G[f"prepass_{name}_in"], G[f"prepass_{name}_out"], G[f"prepass_{name}_freq"]
    got, freq, status = O.hackrf_prepass(tin, tin.shape[1])
    assert np.array_equal(got, want)
    assert np.array_equal(centre(freq), wfreq)
    assert int((status >> 8).sum()) == len(mismatch_lines(G[f"prepass_{name}_text"]))
    assert np.array_equal(status & 1, (tin[:, 0] == 0x7F) & (tin[:, 1] == 0x7F))


@pytest.mark.parametrize("name", CASES)
def test_host_source_prepass_matches_reference(name, tmp_path):
    tin = G[f"prepass_{name}_in"]
    valid = tin.shape[1]
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    tin.tofile(fin)
    r = subprocess.run([SELFTEST, "hackrf_prepass", str(N), str(FS), repr(START), repr(STOP), str(valid), fin, fout],
                       capture_output=True, text=True, timeout=60, env=ENV)
    assert r.returncode == 0, r.stderr
    rec = np.fromfile(fout, np.uint8).reshape(tin.shape[0], 8 + valid)
    assert np.array_equal(rec[:, 8:], G[f"prepass_{name}_out"])
    assert np.array_equal(rec[:, :8].copy().view(np.float64).reshape(-1), G[f"prepass_{name}_freq"])
    assert mismatch_lines(r.stdout) == mismatch_lines(G[f"prepass_{name}_text"])


def expected_messages():
    """What hackRFSource.cpp:224-264 + messageQueue.h:65-91 accept from the golden stream: derived from the
    oracle pre-pass and the reference's bookkeeping rules, (frequency, time, raw chunk) per message."""
    stream = G["sweep_stream"]
    patched, freq, _ = O.hackrf_prepass(stream, VALID)
    table = O.frequency_table(FS, START, STOP)
    index, iteration, current, stamps, marks = 0, 0, 1e12, 0, 0
    msgs = []
    for t in range(stream.shape[0]):
        if iteration >= ITERATIONS:
            break
        c = float(centre(freq[t:t + 1])[0])
        scan_start = False
        if c != current:
            index += 1
            if index >= len(table):
                index, iteration = 0, iteration + 1
            scan_start = index == 0
            current = c
        stamp = 0
        if scan_start:
            stamp = 1500000000 + 1000 * stamps
            stamps += 1
        for k in range(VALID // (2 * N)):
            if stamp:
                marks += 1
            if marks < 2:
                continue
            msgs.append((c, stamp, patched[t, k * 2 * N:(k + 1) * 2 * N]))
    return msgs


def test_host_source_queue_bookkeeping(tmp_path):
    fin, fout = str(tmp_path / "stream.bin"), str(tmp_path / "q.bin")
    G["sweep_stream"].tofile(fin)
    r = subprocess.run([SELFTEST, "hackrf_queue", str(N), str(FS), repr(START), repr(STOP), str(ITERATIONS),
                        str(VALID), fin, fout], capture_output=True, text=True, timeout=60, env=ENV)
    assert r.returncode == 0, r.stderr + r.stdout
    rec = np.fromfile(fout, np.uint8).reshape(-1, 24 + 2 * N)
    want = expected_messages()
    assert rec.shape[0] == len(want) > 0
    f = rec[:, 0:8].copy().view(np.float64).reshape(-1)
    tm = rec[:, 8:16].copy().view(np.int64).reshape(-1)
    seq = rec[:, 16:24].copy().view(np.uint64).reshape(-1)
    assert np.array_equal(seq, np.arange(len(want), dtype=np.uint64))
    assert f.tolist() == [w[0] for w in want]
    assert tm.tolist() == [w[1] for w in want]
    assert all(np.array_equal(rec[i, 24:], want[i][2]) for i in range(len(want)))
    # the reference's stdout says the same: one "Start scan at" per stamped message, hits only from accepted ones
    assert str(G["sweep_text"]).count("Start scan at") == sum(1 for w in want if w[1])


def test_oracle_reproduces_the_reference_sweep_stdout():
    want_hits = GU.parse_hits(str(G["sweep_text"]))
    msgs = expected_messages()
    raw = np.stack([m[2] for m in msgs]).view(np.int8).reshape(-1, N, 2)
    window, use_w = O.window_build(5, N), O.use_window(0.75, N)
    res = O.pipeline(raw, N, FS, 8, 1, True, 1, THR, window, use_w, precision=1)
    got = []
    for b, m in enumerate(msgs):
        _, _, bins = O.detect(res["spectra_db"][b], use_w, 4, THR)
        got += [(O.hit_frequency(m[0], FS, N, int(i)), float(res["spectra_db"][b][(int(i) + N // 2) % N])) for i in bins]
    assert [f for f, _ in got] == [f for f, _ in want_hits] and want_hits
    assert max(abs(a - b) for (_, a), (_, b) in zip(got, want_hits)) < 1e-3


# ---------------------------------------------------------------------------------------------- GPU

@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_device_prepass_matches_reference(name):
    import torch
    from scanner_b200 import SpectrumSense
    tin, want, wfreq = G[f"prepass_{name}_in"], G[f"prepass_{name}_out"], G[f"prepass_{name}_freq"]
    nt, valid = tin.shape
    n = min(N, valid // 2)
    ss = SpectrumSense(sample_count=n, sample_rate=FS, enob=8, sample_kind=1, correct_dc_offset=True,
                       threshold=0.0, window=O.window_build(5, 1024), max_spectra=nt * (valid // (2 * n)))
    d = torch.from_numpy(tin.copy()).cuda()
    freq = torch.zeros(nt, dtype=torch.int64, device="cuda")
    status = torch.zeros(nt, dtype=torch.int32, device="cuda")
    ss.hackrf_prepass_device(d.data_ptr(), nt, valid, freq.data_ptr(), status.data_ptr(),
                             torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(d.cpu().numpy(), want)
    assert np.array_equal(centre(freq.cpu().numpy().view(np.uint64)), wfreq)
    _, ofreq, ostatus = O.hackrf_prepass(tin, valid)
    assert np.array_equal(status.cpu().numpy().view(np.uint32), ostatus)
    # outputs are optional
    d2 = torch.from_numpy(tin.copy()).cuda()
    ss.hackrf_prepass_device(d2.data_ptr(), nt, valid, 0, 0, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(d2.cpu().numpy(), want)


@pytest.mark.gpu
def test_device_prepass_rejects_bad_arguments():
    import torch
    from scanner_b200 import SpectrumSense
    from scanner_b200.binding import ScannerError
    d = torch.zeros(32768, dtype=torch.uint8, device="cuda")
    ss = SpectrumSense(sample_count=1024, sample_rate=FS, enob=8, sample_kind=1, correct_dc_offset=True, threshold=0.0,
                       window=O.window_build(5, 1024))
    with pytest.raises(ScannerError):
        ss.hackrf_prepass_device(d.data_ptr(), 1, 1000)           # not a whole number of buffers
    s16 = SpectrumSense(sample_count=1024, sample_rate=FS, enob=12, sample_kind=3, correct_dc_offset=False, threshold=0.0,
                        window=O.window_build(5, 1024))
    with pytest.raises(ScannerError):
        s16.hackrf_prepass_device(d.data_ptr(), 1, 32768)         # HackRF streams are int8 IQ


@pytest.mark.gpu
def test_device_resident_sweep_stream_matches_oracle():
    """Capture already in HBM: pre-pass kernel + fused kernel over the same bytes, no host touch of samples."""
    import torch
    from scanner_b200 import SpectrumSense
    stream = G["sweep_stream"]
    nt = stream.shape[0]
    chunks = VALID // (2 * N)
    window, use_w = O.window_build(5, N), O.use_window(0.75, N)
    ss = SpectrumSense(sample_count=N, sample_rate=FS, enob=8, sample_kind=1, correct_dc_offset=True,
                       threshold=THR, window=window, max_spectra=nt * chunks)
    d = torch.from_numpy(stream.copy()).cuda()
    freq = torch.zeros(nt, dtype=torch.int64, device="cuda")
    masks = torch.zeros(nt * chunks, ss.words, dtype=torch.int32, device="cuda")
    counts = torch.zeros(nt * chunks, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    ss.hackrf_prepass_device(d.data_ptr(), nt, VALID, freq.data_ptr(), 0, st)
    ss.launch_device(d.data_ptr(), nt * chunks, 0, masks.data_ptr(), counts.data_ptr(), 0, 0, st)
    torch.cuda.synchronize()
    patched, ofreq, _ = O.hackrf_prepass(stream, VALID)
    assert np.array_equal(freq.cpu().numpy().view(np.uint64), ofreq)
    res = O.pipeline(patched.view(np.int8).reshape(-1, N, 2), N, FS, 8, 1, True, 1, THR, window, use_w, precision=1)
    assert np.array_equal(masks.cpu().numpy().view(np.uint32), res["hit_mask"])
    assert np.array_equal(counts.cpu().numpy().view(np.uint32), res["hit_count"])
    assert res["hit_count"].sum() > 0


@pytest.mark.gpu
@pytest.mark.parametrize("threads", [1, 2])
def test_scan_b200_hackrf_prints_what_the_reference_prints(threads, tmp_path):
    f = str(tmp_path / "stream.bin")
    G["sweep_stream"].tofile(f)
    r = subprocess.run([TOOL, "hackrf", str(N), str(FS), repr(START), repr(STOP), repr(THR), str(ITERATIONS),
                        str(VALID), f, str(threads)], capture_output=True, text=True, timeout=120, env=ENV)
    assert r.returncode == 0, r.stderr
    want = str(G["sweep_text"])
    got_hits, want_hits = GU.parse_hits(r.stdout), GU.parse_hits(want)
    assert [h for h, _ in got_hits] == [h for h, _ in want_hits] and want_hits
    assert max(abs(a - b) for (_, a), (_, b) in zip(got_hits, want_hits)) < 1e-3 + 1e-6
    # producer-side lines (the mismatch print) race with the consumer's in both programs: compare per thread
    assert mismatch_lines(r.stdout) == mismatch_lines(want)
    strip = lambda t: [re.sub(r"power_db .*", "", l) for l in t.splitlines()
                       if not l.startswith("interpolateSamples") and not re.match(r"(Starting|Stopped) process thread", l)]
    assert strip(r.stdout) == strip(want)      # table dump, source banner, every "Start scan at <stamp>", every hit
