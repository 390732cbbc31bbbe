"""Synthetic HackRF sweep-mode transfers (what the firmware hands hackRFSource.cpp:224-264): every
16 384-byte block starts with the 10-byte frame header 0x7F 0x7F + little-endian uint64 tuned
frequency; the rest is int8 IQ.  Used by tests/golden/make_golden_hackrf.py and the tests."""
from __future__ import annotations

import numpy as np

from . import synth

BLOCK_BYTES = 16384


def scan_parameters(fs: int, start_hz: float):
    """(step_width, offset) as HackRFSource's constructor stores them, hackRFSource.cpp:108-112."""
    step_width = int(0.75 * fs)                 # uint32_t m_scanStepWidth = 0.75 * sampleRate
    offset = int(step_width / 2.0)              # uint32_t m_scanOffset
    return step_width, offset


def header(freq_hz: int) -> np.ndarray:
    return np.frombuffer(b"\x7f\x7f" + int(freq_hz).to_bytes(8, "little"), np.uint8)


def make_stream(n: int, fs: int, start_hz: float, steps: int, transfers_per_step: int, n_transfers: int,
                valid_length: int, seed: int) -> np.ndarray:
    """uint8 [n_transfers][valid_length]; step k is tuned to start + k * step_width."""
    step_width, _ = scan_parameters(fs, start_hz)
    chunks = valid_length // (2 * n)
    raw = synth.make_buffers(synth.KIND_BYTE_COMPLEX, n, n_transfers * chunks, 8, seed)
    out = raw.reshape(n_transfers, valid_length).view(np.uint8).copy()
    for t in range(n_transfers):
        k = (t // transfers_per_step) % steps
        f = int(start_hz) + k * step_width
        for off in range(0, valid_length, BLOCK_BYTES):
            out[t, off:off + 10] = header(f)
    return out


def prepass_cases(seed: int = 77):
    """Edge cases of interpolateSamples: [(name, valid_length, uint8 [T][valid_length])]."""
    rng = np.random.default_rng(synth.SEED0 + seed)
    cases = []

    def noise(t, valid):
        # random bytes only where interpolateSamples looks (the head of the transfer and the sample before each
        # 8192-sample boundary); zeros elsewhere keep the committed fixture small
        out = np.zeros((t, valid), np.uint8)
        for lo in [0] + list(range(BLOCK_BYTES - 32, valid, BLOCK_BYTES)):
            hi = min(valid, lo + 64)
            out[:, lo:hi] = rng.integers(-100, 101, size=(t, hi - lo), dtype=np.int64).astype(np.int8).view(np.uint8)
        return out

    a = noise(4, 32768)
    for t in range(4):
        for off in range(0, 32768, BLOCK_BYTES):
            a[t, off:off + 10] = header(2_400_000_000 + 15_000_000 * t)
    a[1, 10:12] = np.array([-7, -100], np.int8).view(np.uint8)          # negative patch values
    cases.append(("headers", 32768, a))

    b = noise(5, 65536)
    for t in range(5):
        b[t, 0:10] = header(88_000_000 + t)
        b[t, 10:12] = 0x7F                                               # sample 5 saturated: later iterations act
    b[0, 2 * 8191:2 * 8191 + 2] = 0x7F                                   # average stays 0x7F -> acts again at i = 16384
    b[0, 2 * 16383:2 * 16383 + 2] = np.array([-3, -128], np.int8).view(np.uint8)
    b[1, 2 * 8191:2 * 8191 + 2] = np.array([-128, 126], np.int8).view(np.uint8)   # (127-128)/2 truncates to 0
    b[2, 2 * 8191:2 * 8191 + 2] = np.array([126, 127], np.int8).view(np.uint8)
    b[3, 2 * 8191:2 * 8191 + 2] = 0x7F
    b[3, 2 * 16383:2 * 16383 + 2] = 0x7F
    b[3, 2 * 24575:2 * 24575 + 2] = 0x7F
    b[4, 2:10] = 0x7F                                                    # header frequency equals the patched pattern
    b[4, 2 * 8191:2 * 8191 + 2] = 0x7F
    cases.append(("saturated", 65536, b))

    c = noise(4, 16384)
    c[0, 0:2] = np.array([0x7F, 0x7E], np.uint8)                         # no frame marker
    c[1, 0:2] = np.array([0x00, 0x7F], np.uint8)
    c[2, 0:10] = header(0)                                               # marker with frequency 0
    c[3, 0:10] = header(0xFFFFFFFFFFFFFFFF)
    cases.append(("no_header", 16384, c))

    d = noise(2, 262144)                                                 # libhackrf's real transfer size: 16 iterations
    for t in range(2):
        for off in range(0, 262144, BLOCK_BYTES):
            d[t, off:off + 10] = header(5_990_000_000 + t)
    d[1, 10:12] = 0x7F
    for i in range(8192, 131072, 8192):
        d[1, 2 * (i - 1):2 * (i - 1) + 2] = 0x7F if i < 5 * 8192 else 0x10
    cases.append(("full_transfer", 262144, d))

    e = noise(3, 4096)                                                   # shorter than one block: one iteration
    e[0, 0:10] = header(433_920_000)
    e[2, 0:10] = header(1)
    cases.append(("short", 4096, e))
    return cases
