#!/usr/bin/env python
"""Test infrastructure (lives under tests/ because it uses the oracle): one small run of the cluster kernel through the host
path against the oracle, printable and short enough to run under compute-sanitizer.  usage: cluster_debug.py <kind> <log2n> <dc> <K>"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O
import scanner_b200 as S
from tests import synth
kind, log2n, dc, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
n = 1 << log2n
enob = {1: 8, 2: 12, 3: 12, 4: 0}[kind]
ns = 3
raw = synth.make_buffers(kind, n, ns * K, enob, seed=5)
window = S.window_build(5, n)
use_w = S.use_window(0.75, n)
truth = O.pipeline(raw, n, 8_000_000, enob, kind, bool(dc), K, 0.0, window, use_w, precision=1, want_f64=True)
thr = synth.guard_banded_threshold(truth["spectra_db64"], n, use_w)
truth = O.pipeline(raw, n, 8_000_000, enob, kind, bool(dc), K, thr, window, use_w, precision=1, want_f64=True)
with S.SpectrumSense(n, 8_000_000, enob, thr, window, sample_kind=kind, correct_dc_offset=bool(dc), averaging=K, max_spectra=ns) as ss:
    print(ss.kernel_name)
    got = ss.process(raw)
err = np.abs(got["spectra_db"].astype(np.float64) - truth["spectra_db64"])
strong = truth["spectra_db64"] > np.median(truth["spectra_db64"])
print("max dB err on strong bins", err[strong].max(), "masks equal", np.array_equal(got["hit_mask"], truth["hit_mask"]),
      "counts", got["hit_count"], truth["hit_count"])
hits_ok = all(np.array_equal(got["hits"]["bin"][s, :truth["hit_count"][s]], np.nonzero(np.unpackbits(truth["hit_mask"][s].view(np.uint8), bitorder="little"))[0]) for s in range(ns))
print("hit records ordered and complete:", hits_ok)
