"""Generates tests/golden/record_vectors.npz: the reference's own trigger/record path (process.cpp:239-270,
303-313 + the writer thread of messageQueue.h:98-139, compiled unmodified into oracle/_ref/ref_tool) on a
scenario with loud buffers that trip `triggerCount > 1047` (process.cpp:62).  Stored: the raw buffers, the
reference's stdout and, per recorded file, its name suffix, size and SHA-256 (the content is the converted
samples of a window of messages, which the tests rebuild from the converter oracle).
Run in the build container:  python tests/golden/make_golden_record.py"""
import hashlib
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle as O   # noqa: E402
from tests import synth   # noqa: E402

REF_TOOL = os.path.join(ROOT, "oracle", "_ref", "ref_tool")
ENV = dict(os.environ, TZ="UTC")


def main():
    n, fs, enob, kind, dc = 2048, 20_000_000, 8, 1, True
    per_sweep, pre, post = 4, 2, 3
    n_buf = 36
    loud = [9, 10, 11, 20, 29]                                     # buffer indices (sequence id = index - per_sweep)
    rng = np.random.default_rng(0x5CA77E2 + 31)
    raw = np.empty((n_buf, n, 2), np.int8)
    for b in range(n_buf):
        sigma = 40.0 if b in loud else 1.5
        raw[b] = np.clip(np.rint(rng.standard_normal((n, 2)) * sigma + 1.0), -128, 127)
    freqs = 2.4e9 + 15e6 * (np.arange(n_buf) % per_sweep)
    window, use_w = O.window_build(5, n), O.use_window(0.75, n)
    truth = O.pipeline(raw, n, fs, enob, kind, dc, 1, 0.0, window, use_w, precision=1, want_f64=True)
    cand = truth["spectra_db64"][:, synth.candidate_bins(n, use_w)]
    quiet_max = max(cand[b].max() for b in range(n_buf) if b not in loud)
    loud_counts_at = lambda t: [int((cand[b] > t).sum()) for b in loud]
    thr = float(np.float32(quiet_max + 1.0))                      # no quiet bin within 1 dB: quiet buffers print nothing
    assert min(loud_counts_at(thr)) > 1200, loud_counts_at(thr)   # comfortably above 1047
    with tempfile.TemporaryDirectory() as d:
        rp, fp = os.path.join(d, "raw.bin"), os.path.join(d, "freq.bin")
        raw.tofile(rp)
        freqs.astype(np.float64).tofile(fp)
        base = os.path.join(d, "rec-")
        r = subprocess.run([REF_TOOL, "record", str(kind), str(n), repr(float(fs)), str(enob), "1", repr(thr), "5",
                            str(per_sweep), rp, fp, base, str(pre), str(post)],
                           capture_output=True, text=True, check=True, env=ENV, timeout=300)
        files = sorted(f for f in os.listdir(d) if f.startswith("rec-"))
        recs = []
        for f in files:
            data = open(os.path.join(d, f), "rb").read()
            recs.append((f[len("rec-"):], len(data), hashlib.sha256(data).hexdigest()))
        text = r.stdout.replace(d + os.sep, "")
    lines = [l for l in text.splitlines() if not l.startswith("freq ")]
    print("\n".join(lines))
    print(recs)
    assert recs
    out = dict(raw=raw, freqs=freqs, text=np.array(text), files=np.array(recs),
               params=np.array([n, fs, enob, kind, int(dc), per_sweep, pre, post, thr], np.float64))
    path = os.path.join(ROOT, "tests", "golden", "record_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
