#!/usr/bin/env python
"""Generates tests/golden/reference_vectors.npz by running the reference's OWN sources
(oracle/_ref/ref_tool: /root/reference/{utility,fft,process,frequencyTable}.cpp + messageQueue.h
compiled unmodified against the shim headers in oracle/shim) on seeded inputs.

Run in the build container (needs /root/reference):  make -C oracle ref && python tests/golden/make_golden.py
The .npz is committed; tests only read it, so they run without the reference tree."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle as O            # noqa: E402
from oracle import ref as R   # noqa: E402
from tests import synth       # noqa: E402

assert R.available(), "build oracle/_ref first: make -C oracle ref"
out = {}
rng = np.random.default_rng(20261017)

# ---- converters (utility.cpp:9-84), incl. the probe vectors of SURVEY.md section 8c KAT-1 -----------
conv_cases = []
def add_conv(name, kind, n, enob, dc, raw):
    got = R.convert(kind, raw, n, enob, dc)
    out[f"conv_{name}_raw"] = raw
    out[f"conv_{name}_out"] = got
    conv_cases.append((name, kind, n, enob, int(dc)))

add_conv("i8_probe", 1, 1, 8, False, np.array([[[10, 20]]], np.int8))
add_conv("i16_probe", 3, 1, 12, False, np.array([[[100, 200]]], np.int16))
add_conv("i16_negsum_quirk", 3, 4, 12, True, np.array([[[-100, -200]] * 4], np.int16))
add_conv("i16_enob16_signflip", 3, 4, 16, False, np.array([[[100, -200], [32767, -32768], [1, -1], [0, 5]]], np.int16))
add_conv("i8_dc_mean", 1, 4, 8, True, np.array([[[10, 20], [12, 22], [11, 21], [11, 21]]], np.int8))
add_conv("i8_dc_negsum_quirk", 1, 256, 8, True, rng.integers(-128, -100, (2, 256, 2)).astype(np.int8))
add_conv("i8_rand", 1, 256, 8, False, rng.integers(-128, 128, (3, 256, 2)).astype(np.int8))
add_conv("i8_rand_dc", 1, 256, 8, True, rng.integers(-60, 128, (3, 256, 2)).astype(np.int8))
add_conv("i8_enob6", 1, 64, 6, True, rng.integers(-32, 32, (2, 64, 2)).astype(np.int8))
add_conv("i16_rand", 3, 512, 12, False, rng.integers(-2048, 2048, (2, 512, 2)).astype(np.int16))
add_conv("i16_rand_dc", 3, 512, 12, True, rng.integers(-1000, 2048, (2, 512, 2)).astype(np.int16))
add_conv("i16_split", 2, 512, 12, False, rng.integers(-2048, 2048, (2, 2, 512)).astype(np.int16))
add_conv("i16_split_dc", 2, 512, 14, True, rng.integers(-8192, 8192, (2, 2, 512)).astype(np.int16))
out["conv_cases"] = np.array(conv_cases, dtype="U32")

# ---- dB (utility.cpp:86-98), both header readings -------------------------------------------------
x = np.concatenate([
    np.array([0, 1, 3 + 4j, 1e-3 + 2e-3j, 1e-20 + 0j, 1e18 + 1e18j, -2.5 + 0.5j], np.complex64),
    ((rng.standard_normal(4096) + 1j * rng.standard_normal(4096)) * 10.0 ** rng.uniform(-6, 6, 4096)).astype(np.complex64),
])
out["mag_in"] = x
out["mag_out_math"] = R.magnitude(x, "math")
out["mag_out_cmath"] = R.magnitude(x, "cmath")

# ---- frequency tables (frequencyTable.cpp:9-37) -------------------------------------------------
ft_cases = [(20_000_000, 2.4e9, 3.15e9, 0.75, 0.0), (56_000_000, 1e9, 2e9, 0.75, 0.0),
            (10_000_000, 0.1e9, 1.1e9, 0.75, 0.0), (8_000_000, 3e8, 0.0, 0.75, 0.0),
            (8_000_000, 88e6, 108e6, 0.75, 0.05), (2_400_000, 24e6, 1.7e9, 0.75, 0.0)]
out["ft_cases"] = np.array(ft_cases, np.float64)
for i, c in enumerate(ft_cases):
    out[f"ft_{i}"] = R.frequency_table(int(c[0]), c[1], c[2], c[3], c[4])

# ---- whole pipeline through the reference's SampleQueue + ProcessSamples ---------------------------
scan_cases = []
def add_scan(name, kind, n, fs, enob, dc, per_sweep, sweeps, seed, win=5, mode=2, quantile=0.985):
    nbuf = per_sweep * sweeps
    raw = synth.make_buffers(kind, n, nbuf, enob, seed)
    table = O.frequency_table(fs, 400e6, 400e6 + per_sweep * 0.75 * fs - 1)
    assert len(table) == per_sweep, (len(table), per_sweep)
    freqs = np.tile(table, sweeps)
    window = O.window_build(win, n)
    use_w = O.use_window(0.75, n)
    if mode == 2:
        truth = O.pipeline(raw, n, fs, enob, kind, dc, 1, 0.0, window, use_w, precision=1, want_f64=True)
        thr = synth.guard_banded_threshold(truth["spectra_db64"], n, use_w, quantile=quantile)
    else:
        # normalised samples have |x| < 1, so their "dB" is negative and maxMagnitude stays at its
        # seed numeric_limits<float>::min() (process.cpp:207); drive every other buffer into
        # saturation so that some buffers trigger and some do not.
        for b in range(1, nbuf, 2):
            if kind == 2:
                raw[b, :, :8] = raw.max()
            else:
                raw[b, :8, :] = np.iinfo(raw.dtype).min if b % 4 == 1 else np.iinfo(raw.dtype).max
        _, mm = O.time_domain(raw, n, enob, kind, dc, 0.0)
        hi = np.sort(mm[1::2, 0])
        thr = float(np.float32(0.5 * hi[0]))
        assert thr > 1e-3
    text = R.scan(kind, raw, freqs, n, fs, enob, dc, thr, win, mode, per_sweep)
    out[f"scan_{name}_raw"] = raw
    out[f"scan_{name}_freqs"] = freqs
    out[f"scan_{name}_text"] = np.array(text)
    scan_cases.append((name, kind, n, fs, enob, int(dc), per_sweep, sweeps, win, mode, repr(thr)))

add_scan("i8_dc_256", 1, 256, 20_000_000, 8, True, 4, 3, seed=11)
add_scan("i8_1024", 1, 1024, 20_000_000, 8, False, 3, 3, seed=12)
add_scan("i16_dc_512", 3, 512, 8_000_000, 12, True, 3, 3, seed=13)
add_scan("i16split_256", 2, 256, 8_000_000, 12, False, 3, 2, seed=14)
add_scan("f32_hann_1024", 4, 1024, 10_000_000, 0, False, 3, 3, seed=15, win=1)
add_scan("i8_dc_2048", 1, 2048, 20_000_000, 8, True, 2, 3, seed=16)
add_scan("td_i16_512", 3, 512, 8_000_000, 12, True, 4, 3, seed=17, mode=1)
add_scan("td_i8_256", 1, 256, 20_000_000, 8, False, 4, 3, seed=18, mode=1)
out["scan_cases"] = np.array(scan_cases, dtype="U40")

path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_vectors.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path), "bytes;", len(conv_cases), "converter,", len(ft_cases),
      "frequency-table,", len(scan_cases), "scan cases")
