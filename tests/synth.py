"""Seeded synthetic IQ for the parity tests: complex white Gaussian
noise (sigma = 0.05 FS per rail) + up to 4 complex tones at random shifted-bin centres inside
the used band, amplitudes log-uniform in [-40, -3] dBFS, a small positive DC, quantised with
round-to-nearest and saturation.  Also the guard-banded threshold chooser."""
from __future__ import annotations

import numpy as np

KIND_BYTE_COMPLEX, KIND_SHORT, KIND_SHORT_COMPLEX, KIND_FLOAT_COMPLEX = 1, 2, 3, 4
SEED0 = 0x5CA77E2


def full_scale(kind: int, enob: int) -> float:
    if kind == KIND_FLOAT_COMPLEX:
        return 1.0
    return float((1 << (enob - 1)) - 1)


def make_buffers(kind: int, n: int, n_buffers: int, enob: int, seed: int, tones_max: int = 4,
                 sigma: float = 0.05, dc: float = 0.01, use_bw: float = 0.75) -> np.ndarray:
    """Returns the raw array: int8 [B][N][2], int16 [B][N][2], int16 split [B][2][N] or
    float32 [B][N][2]."""
    rng = np.random.default_rng(SEED0 + seed)
    fs_amp = full_scale(kind, enob)
    t = np.arange(n)
    use_w = int(use_bw * n / 2.0)
    out = np.empty((n_buffers, n), np.complex128)
    for b in range(n_buffers):
        x = sigma * (rng.standard_normal(n) + 1j * rng.standard_normal(n)) + dc * (1 + 1j)
        for _ in range(int(rng.integers(0, tones_max + 1))):
            i = int(rng.integers(n // 2 - use_w, n // 2 + use_w + 1))   # shifted bin
            k = (i + n // 2) % n                                         # FFT bin
            amp = 10.0 ** (rng.uniform(-40.0, -3.0) / 20.0)
            x = x + amp * np.exp(1j * (2 * np.pi * k * t / n + rng.uniform(0, 2 * np.pi)))
        out[b] = x
    if kind == KIND_FLOAT_COMPLEX:
        raw = np.empty((n_buffers, n, 2), np.float32)
        raw[..., 0] = out.real
        raw[..., 1] = out.imag
        return raw
    lo, hi = (-128, 127) if kind == KIND_BYTE_COMPLEX else (-32768, 32767)
    re = np.clip(np.rint(out.real * fs_amp), max(lo, -fs_amp - 1), min(hi, fs_amp))
    im = np.clip(np.rint(out.imag * fs_amp), max(lo, -fs_amp - 1), min(hi, fs_amp))
    dt = np.int8 if kind == KIND_BYTE_COMPLEX else np.int16
    if kind == KIND_SHORT:
        raw = np.empty((n_buffers, 2, n), dt)
        raw[:, 0, :] = re
        raw[:, 1, :] = im
        return raw
    raw = np.empty((n_buffers, n, 2), dt)
    raw[..., 0] = re
    raw[..., 1] = im
    return raw


def candidate_bins(n: int, use_w: int, dc_w: int = 4) -> np.ndarray:
    """FFT-bin indices j that process.cpp:46-53 does not skip."""
    i = np.arange(n, dtype=np.int64)
    j = (i + n // 2) % n
    keep = ~((j < dc_w) | ((n - j) < dc_w))
    keep &= ~((i < (n // 2 - use_w)) | (i > (n // 2 + use_w)))
    return j[keep]


def guard_banded_threshold(db64: np.ndarray, n: int, use_w: int, guard: float = 5e-3,
                           quantile: float = 0.97) -> float:
    """A threshold near the given quantile of the candidate-bin dB values (so there are both
    hits and misses) that no candidate bin of any spectrum comes within +-guard dB of."""
    cand = db64[:, candidate_bins(n, use_w)].ravel()
    cand = np.sort(cand[np.isfinite(cand)])
    start = int(quantile * (cand.size - 1))
    for k in range(start, cand.size - 1):
        gap = cand[k + 1] - cand[k]
        if gap > 4 * guard:
            thr = np.float32(0.5 * (cand[k] + cand[k + 1]))
            if np.min(np.abs(cand - float(thr))) > guard:
                return float(thr)
    raise AssertionError("no guard-banded threshold found")
