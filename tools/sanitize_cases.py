"""Small invocations of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scanner_b200 as S
from tests import synth
cases = [(1, 2048, 8, True, 1), (1, 2048, 8, False, 1),          # WPT
         (1, 8192, 8, True, 1), (3, 4096, 12, True, 3), (4, 8192, 0, False, 1), (4, 4096, 0, False, 2),   # P64
         (3, 1024, 12, True, 2), (2, 512, 12, True, 1), (1, 256, 8, True, 1), (4, 16384, 0, False, 1),    # generic
         (1, 32768, 8, True, 1)]                                                                              # four-step
for kind, n, enob, dc, K in cases:
    raw = synth.make_buffers(kind, n, 5 * K, enob, seed=n + kind)
    w = S.window_build(5, n)
    with S.SpectrumSense(n, 8_000_000, enob, 8.0, w, sample_kind=kind, correct_dc_offset=dc, averaging=K,
                         max_spectra=5, max_hits_per_spectrum=64) as ss:
        r = ss.process(raw)
        print(ss.kernel_name, int(r["hit_count"].sum()), flush=True)
raw = synth.make_buffers(1, 4096, 6, 8, seed=3)
with S.SpectrumSense(4096, 8_000_000, 8, -5.0, None, sample_kind=1, correct_dc_offset=True, mode=S.MODE_TIME_DOMAIN, max_spectra=6) as ss:
    print("time domain", ss.process(raw)["hit_count"].tolist())
