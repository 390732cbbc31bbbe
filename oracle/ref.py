"""Runs oracle/_ref/ref_tool (the reference's own sources built with shim third-party headers).
TEST INFRASTRUCTURE; only present where `make -C oracle ref` could see /root/reference (the built
binaries travel to the GPU box, the reference tree does not)."""
from __future__ import annotations

import os
import re
import subprocess
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def tool(variant: str = "math") -> str:
    return os.path.join(_HERE, "_ref", "ref_tool" if variant == "math" else "ref_tool_cmath")


def available() -> bool:
    return os.access(tool(), os.X_OK) and os.access(tool("cmath"), os.X_OK)


def convert(kind: int, raw: np.ndarray, n: int, enob: int, dc: bool) -> np.ndarray:
    out = subprocess.run([tool(), "convert", str(kind), str(n), str(enob), "1" if dc else "0"],
                         input=np.ascontiguousarray(raw).tobytes(), stdout=subprocess.PIPE, check=True).stdout
    return np.frombuffer(out, np.float32).reshape(-1, n, 2)


def magnitude(fft: np.ndarray, variant: str = "math") -> np.ndarray:
    fft = np.ascontiguousarray(fft, np.complex64)
    out = subprocess.run([tool(variant), "magnitude", str(fft.shape[0])], input=fft.tobytes(),
                         stdout=subprocess.PIPE, check=True).stdout
    return np.frombuffer(out, np.float32).copy()


def frequency_table(fs: int, start: float, stop: float, use_bw: float = 0.75, dc_ignore: float = 0.0) -> np.ndarray:
    out = subprocess.run([tool(), "freqtable", repr(float(fs)), repr(start), repr(stop), repr(use_bw),
                          repr(dc_ignore)], stdout=subprocess.PIPE, check=True).stdout.decode()
    return np.array([float(m.group(1)) for m in re.finditer(r"Frequency \d+: (-?\d+)", out)], np.float64)


def scan(kind: int, raw: np.ndarray, freqs: np.ndarray, n: int, fs: int, enob: int, dc: bool, threshold: float,
         win_type: int = 5, mode: int = 2, buffers_per_sweep: int = 0) -> str:
    """The reference's stdout for these buffers (one ProcessSamples worker)."""
    with tempfile.TemporaryDirectory() as d:
        rp, fp = os.path.join(d, "raw.bin"), os.path.join(d, "freq.bin")
        np.ascontiguousarray(raw).tofile(rp)
        np.ascontiguousarray(freqs, np.float64).tofile(fp)
        out = subprocess.run([tool(), "scan", str(kind), str(n), repr(float(fs)), str(enob), "1" if dc else "0",
                              repr(float(threshold)), str(win_type), str(mode), str(buffers_per_sweep), rp, fp],
                             stdout=subprocess.PIPE, check=True).stdout.decode()
    return out


def parse_hits(text: str):
    """[(freq_hz, power_db)] from 'freq %lu power_db %f' lines, in print order."""
    return [(int(m.group(1)), float(m.group(2))) for m in re.finditer(r"freq (\d+) power_db (-?[\d.]+|-?inf|-?nan)", text)]
