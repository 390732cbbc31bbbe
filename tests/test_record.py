"""Trigger/record path: `triggerCount > 1047` (process.cpp:62), pre/post-trigger
window bookkeeping (process.cpp:239-270, 311-313) and the recording files (messageQueue.h:98-139).

Golden: tests/golden/record_vectors.npz = stdout and per-file SHA-256 of the reference's own code on a scenario
with loud buffers (tests/golden/make_golden_record.py).  The files hold the converted samples (fftwf_complex) of a
window of queue messages; here the queue holds raw int8, so the GPU path converts on write (scn_convert_host)."""
import hashlib
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle as O   # noqa: E402

G = np.load(os.path.join(ROOT, "tests", "golden", "record_vectors.npz"), allow_pickle=False)
TOOL = os.path.join(ROOT, "scanner_b200", "scan_b200")
ENV = dict(os.environ, TZ="UTC")
N, FS, ENOB, KIND, DC, PER_SWEEP, PRE, POST = [int(x) for x in G["params"][:8]]
THR = float(G["params"][8])
FILES = [(str(a), int(b), str(c)) for a, b, c in G["files"]]


def record_lines(text):
    return [l for l in str(text).splitlines() if re.match(r"BeginWrite|EndWrite|Writing", l)]


def windows_from_text(text):
    """[(start, end)] in sequence ids from the reference's BeginWrite / EndWrite lines."""
    starts = [int(m.group(1)) for m in re.finditer(r"BeginWrite \S+: (\d+)", str(text))]
    ends = [int(m.group(1)) for m in re.finditer(r"EndWrite (\d+)", str(text))]
    return list(zip(starts, ends))


def test_reference_files_are_the_converted_window_messages():
    """Pins what a recording IS (and the converter oracle once more): file k == converted messages [start_k, end_k)."""
    raw = G["raw"]
    wins = windows_from_text(G["text"])
    assert len(wins) == len(FILES) == 3
    for (start, end), (_, size, sha) in zip(wins, FILES):
        msgs = [O.convert(KIND, raw[PER_SWEEP + s], N, ENOB, bool(DC)) for s in range(start, end)]   # first sweep dropped
        blob = b"".join(np.ascontiguousarray(m, np.float32).tobytes() for m in msgs)
        assert len(blob) == size
        assert hashlib.sha256(blob).hexdigest() == sha


def test_trigger_rule_matches_the_reference_windows():
    """process.cpp:62,250-270: a buffer with more than 1047 hits triggers; window = [first - pre, last + post + 1)."""
    window, use_w = O.window_build(5, N), O.use_window(0.75, N)
    res = O.pipeline(G["raw"], N, FS, ENOB, KIND, bool(DC), 1, THR, window, use_w, precision=1)
    trig = [b - PER_SWEEP for b in range(PER_SWEEP, G["raw"].shape[0]) if res["hit_count"][b] > 1047]
    wins, writing, end = [], False, 0
    for s in range(G["raw"].shape[0] - PER_SWEEP):
        t = s in trig
        if writing:
            if t:
                end = max(end, s + POST + 1)
            elif s == end:
                wins[-1] = (wins[-1][0], s)
                writing = False
        elif t:
            wins.append((s - min(s, PRE), None))
            writing, end = True, s + POST + 1
    assert wins == windows_from_text(G["text"])


@pytest.mark.gpu
@pytest.mark.parametrize("threads", [1])
def test_scan_b200_records_what_the_reference_records(threads, tmp_path):
    rp, fp = str(tmp_path / "raw.bin"), str(tmp_path / "freq.bin")
    G["raw"].tofile(rp)
    G["freqs"].astype(np.float64).tofile(fp)
    base = str(tmp_path / "rec-")
    r = subprocess.run([TOOL, "record", str(KIND), str(N), repr(float(FS)), str(ENOB), str(DC), repr(THR), "5",
                        str(PER_SWEEP), rp, fp, base, str(PRE), str(POST), str(threads)],
                       capture_output=True, text=True, timeout=120, env=ENV)
    assert r.returncode == 0, r.stderr
    got = r.stdout.replace(str(tmp_path) + os.sep, "")
    # the writer thread prints asynchronously in both programs: compare its lines and the worker's separately
    pick = lambda t, pat: [l for l in str(t).splitlines() if re.match(pat, l)]
    assert pick(got, r"BeginWrite|EndWrite") == pick(G["text"], r"BeginWrite|EndWrite")
    assert pick(got, r"Writing") == pick(G["text"], r"Writing")
    files = sorted(f for f in os.listdir(tmp_path) if f.startswith("rec-"))
    assert [f[len("rec-"):] for f in files] == [name for name, _, _ in FILES]
    for f, (_, size, sha) in zip(files, FILES):
        data = open(os.path.join(tmp_path, f), "rb").read()
        assert len(data) == size
        assert hashlib.sha256(data).hexdigest() == sha        # bit for bit the reference's recording
