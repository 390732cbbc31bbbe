"""Generates tests/golden/convert_vectors.npz: the reference's own converters (utility.cpp:9-84, compiled
unmodified into oracle/_ref/ref_tool) on buffers of the sizes the C ABI accepts (N >= 256), so the standalone
GPU converter (scn_convert_*) can be checked bit for bit.  Inputs are stored sparsely-random to keep the file
small.  Run in the build container:  python tests/golden/make_golden_convert.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref as R   # noqa: E402

rng = np.random.default_rng(0x5CA77E2 + 9)
out, cases = {}, []


def add(name, kind, n, enob, dc, raw):
    out[f"{name}_raw"] = raw
    out[f"{name}_out"] = R.convert(kind, raw, n, enob, dc)
    cases.append((name, kind, n, enob, int(dc)))


def i8(lo, hi, b, n):
    return rng.integers(lo, hi, (b, n, 2)).astype(np.int8)


def i16(lo, hi, shape):
    return rng.integers(lo, hi, shape).astype(np.int16)


add("i8_256", 1, 256, 8, False, i8(-128, 128, 3, 256))
add("i8_256_dc", 1, 256, 8, True, i8(-40, 128, 3, 256))
add("i8_2048_dc_negsum_quirk", 1, 2048, 8, True, i8(-128, 20, 2, 2048))          # negative sums: unsigned division
add("i8_1024_dc_mixed_sign_sums", 1, 1024, 8, True, np.stack([i8(-128, 128, 1, 1024)[0] + np.array([3, -3], np.int8),
                                                               i8(-100, 100, 1, 1024)[0] - np.array([5, -5], np.int8)]))
add("i8_512_enob6_dc", 1, 512, 6, True, i8(-32, 32, 2, 512))
add("i8_256_saturated", 1, 256, 8, True, np.where(rng.random((2, 256, 2)) < 0.5, 127, -128).astype(np.int8))
add("i16_512", 3, 512, 12, False, i16(-2048, 2048, (2, 512, 2)))
add("i16_512_dc", 3, 512, 12, True, i16(-1000, 2048, (2, 512, 2)))
add("i16_256_dc_negsum_quirk", 3, 256, 12, True, i16(-2048, 100, (2, 256, 2)))
add("i16_256_enob16_signflip", 3, 256, 16, True, i16(-32768, 32768, (2, 256, 2)))
add("i16_4096_enob14_dc", 3, 4096, 14, True, i16(-8192, 8192, (1, 4096, 2)))
add("i16s_512", 2, 512, 12, False, i16(-2048, 2048, (2, 2, 512)))
add("i16s_512_dc", 2, 512, 14, True, i16(-8192, 8192, (2, 2, 512)))
add("i16s_1024_dc_negsum_quirk", 2, 1024, 12, True, i16(-2048, 50, (2, 2, 1024)))
out["cases"] = np.array(cases)
path = os.path.join(ROOT, "tests", "golden", "convert_vectors.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path), "bytes")
