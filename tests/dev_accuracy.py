#!/usr/bin/env python
"""FFT accuracy of the GPU path vs the double oracle and vs the CPU fp32 restatement (developer tool).
usage: accuracy.py <kind> <log2n>"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O, scanner_b200 as S
from tests import synth
kind, log2n = int(sys.argv[1]), int(sys.argv[2]); n = 1 << log2n
enob = {1: 8, 2: 12, 3: 12, 4: 0}[kind]
raw = synth.make_buffers(kind, n, 8, enob, seed=99)
w = S.window_build(5, n); uw = S.use_window(0.75, n)
t64 = O.pipeline(raw, n, 8_000_000, enob, kind, kind != 4, 1, 0.0, w, uw, precision=1, want_f64=True)["spectra_db64"]
c32 = O.pipeline(raw, n, 8_000_000, enob, kind, kind != 4, 1, 0.0, w, uw, precision=0)["spectra_db"].astype(np.float64)
with S.SpectrumSense(n, 8_000_000, enob, 1e9, w, sample_kind=kind, correct_dc_offset=kind != 4, max_spectra=8) as ss:
    g = ss.process(raw)["spectra_db"].astype(np.float64)
mt = 10 ** (t64 / 10); rms = np.sqrt(np.mean(mt ** 2, axis=1, keepdims=True))
lg = np.abs(10 ** (g / 10) - mt) / rms; lc = np.abs(10 ** (c32 / 10) - mt) / rms
strong = t64 >= 10 * np.log10(rms) - 10
print(f"{os.path.basename(os.environ.get('SCN_LIB','default')):20s} kind={kind} N=2^{log2n}: GPU lin err rms {np.sqrt(np.mean(lg**2)):.2e} max {lg.max():.2e} | "
      f"CPU fp32 rms {np.sqrt(np.mean(lc**2)):.2e} max {lc.max():.2e} | max dB err on strong bins GPU {np.abs(g-t64)[strong].max():.2e} CPU {np.abs(c32-t64)[strong].max():.2e}")
