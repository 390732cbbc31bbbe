"""ctypes binding of oracle/scanner_oracle.cpp (the checker; never the product path)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

KIND_BYTE_COMPLEX, KIND_SHORT, KIND_SHORT_COMPLEX, KIND_FLOAT_COMPLEX = 1, 2, 3, 4
WIN_HAMMING, WIN_HANN, WIN_BLACKMAN, WIN_RECTANGULAR, WIN_BLACKMAN_HARRIS = 0, 1, 2, 3, 5

_VP, _U32, _F, _D, _I = C.c_void_p, C.c_uint32, C.c_float, C.c_double, C.c_int


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "scanner_oracle.cpp")
    if force or not os.path.exists(_SO) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(_SO)):
        subprocess.check_call(["make", "-C", _HERE, "_build/liboracle.so"], stdout=subprocess.DEVNULL)
    return _SO


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        h = C.CDLL(_SO)
        h.orc_bytes_per_sample.restype = _U32
        h.orc_bytes_per_sample.argtypes = [_U32]
        h.orc_convert.restype = None
        h.orc_convert.argtypes = [_U32, _VP, _VP, _U32, _U32, _U32]
        h.orc_window_build.restype = None
        h.orc_window_build.argtypes = [_I, _U32, _VP]
        h.orc_window_apply.restype = None
        h.orc_window_apply.argtypes = [_VP, _VP, _U32]
        h.orc_fft_f32.restype = None
        h.orc_fft_f32.argtypes = [_VP, _VP, _U32]
        h.orc_fft_f64.restype = None
        h.orc_fft_f32_batch8.argtypes = [_VP, _VP, _U32]
        h.orc_fft_f32_batch8.restype = None
        h.orc_fft_f64.argtypes = [_VP, _VP, _U32]
        h.orc_magnitude_db.restype = None
        h.orc_magnitude_db.argtypes = [_VP, _VP, _U32, _I]
        h.orc_use_window.restype = _U32
        h.orc_use_window.argtypes = [_D, _U32]
        h.orc_detect.restype = _U32
        h.orc_detect.argtypes = [_VP, _U32, _U32, _U32, _F, _VP, _VP, _U32]
        h.orc_hit_frequency.restype = C.c_uint64
        h.orc_hit_frequency.argtypes = [_D, _U32, _U32, _U32]
        h.orc_frequency_table.restype = _U32
        h.orc_frequency_table.argtypes = [_U32, _D, _D, _D, _D, _VP, _U32]
        h.orc_pipeline.restype = None
        h.orc_pipeline.argtypes = [_U32, _U32, _U32, _U32, _U32, _U32, _F, _U32, _U32, _I, _VP, _VP, _U32,
                                   _I, _VP, _VP, _VP, _VP, _U32]
        h.orc_time_domain.restype = None
        h.orc_time_domain.argtypes = [_U32, _U32, _U32, _U32, _F, _VP, _U32, _VP, _VP]
        h.orc_bench.restype = _D
        h.orc_bench.argtypes = [_U32, _U32, _U32, _U32, _U32, _U32, _F, _U32, _U32, _VP, _VP, _U32, _U32,
                                _U32, _I, _VP]
        h.orc_hardware_threads.restype = _U32
        h.orc_hackrf_prepass.argtypes = [_VP, _U32, _U32, _VP, _VP]
        h.orc_hackrf_prepass.restype = None
        h.orc_hardware_threads.argtypes = []
        _lib = h
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(_VP)


def bytes_per_sample(kind: int) -> int:
    return int(lib().orc_bytes_per_sample(kind))


def convert(kind: int, raw: np.ndarray, n: int, enob: int, correct_dc: bool) -> np.ndarray:
    raw = np.ascontiguousarray(raw)
    out = np.empty((n, 2), np.float32)
    lib().orc_convert(kind, _p(raw), _p(out), n, enob, 1 if correct_dc else 0)
    return out


def window_build(win_type: int, n: int) -> np.ndarray:
    out = np.empty(n, np.float32)
    lib().orc_window_build(win_type, n, _p(out))
    return out


def window_apply(iq: np.ndarray, w: np.ndarray) -> np.ndarray:
    out = np.ascontiguousarray(iq, np.float32).copy()
    lib().orc_window_apply(_p(out), _p(np.ascontiguousarray(w, np.float32)), out.shape[0])
    return out


def fft_f32(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, np.complex64)
    out = np.empty_like(x)
    lib().orc_fft_f32(_p(x), _p(out), x.shape[0])
    return out


def fft_f32_batch8(x: np.ndarray) -> np.ndarray:
    """x: [8][n][2] float32 -> eight forward FFTs through the SIMD-batched plan of the timed CPU baseline."""
    x = np.ascontiguousarray(x, np.float32)
    assert x.shape[0] == 8 and x.shape[2] == 2
    out = np.empty_like(x)
    lib().orc_fft_f32_batch8(_p(x), _p(out), x.shape[1])
    return out


def fft_f64(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, np.complex128)
    out = np.empty_like(x)
    lib().orc_fft_f64(_p(x), _p(out), x.shape[0])
    return out


def magnitude_db(fft: np.ndarray, variant: int = 0) -> np.ndarray:
    fft = np.ascontiguousarray(fft, np.complex64)
    out = np.empty(fft.shape[0], np.float32)
    lib().orc_magnitude_db(_p(fft), _p(out), fft.shape[0], variant)
    return out


def use_window(use_bandwidth: float, n: int) -> int:
    return int(lib().orc_use_window(use_bandwidth, n))


def detect(db: np.ndarray, use_window_bins: int, dc_ignore: int, threshold: float):
    db = np.ascontiguousarray(db, np.float32)
    n = db.shape[0]
    mask = np.zeros((n + 31) // 32, np.uint32)
    bins = np.zeros(n, np.uint32)
    cnt = lib().orc_detect(_p(db), n, use_window_bins, dc_ignore, threshold, _p(mask), _p(bins), n)
    return int(cnt), mask, bins[:cnt].copy()


def hit_frequency(center: float, sample_rate: int, n: int, i: int) -> int:
    return int(lib().orc_hit_frequency(center, sample_rate, n, i))


def frequency_table(sample_rate: int, start: float, stop: float, use_bw: float = 0.75,
                    dc_ignore: float = 0.0) -> np.ndarray:
    cnt = lib().orc_frequency_table(sample_rate, start, stop, use_bw, dc_ignore, None, 0)
    out = np.zeros(cnt, np.float64)
    lib().orc_frequency_table(sample_rate, start, stop, use_bw, dc_ignore, _p(out), cnt)
    return out


def pipeline(raw: np.ndarray, n: int, sample_rate: int, enob: int, kind: int, correct_dc: bool,
             averaging: int, threshold: float, window: np.ndarray, use_window_bins: int,
             dc_ignore_window: int = 4, precision: int = 1, db_variant: int = 0, threads: int = 1,
             want_f64: bool = False) -> dict:
    """precision 0: fp32 FFT (reference-like); 1: fp64 FFT of the same fp32 windowed samples."""
    raw = np.ascontiguousarray(raw)
    K = max(1, averaging)
    n_spectra = raw.nbytes // (n * bytes_per_sample(kind) * K)
    window = None if window is None else np.ascontiguousarray(window, np.float32)
    spectra = np.empty((n_spectra, n), np.float32)
    spectra64 = np.empty((n_spectra, n), np.float64) if (want_f64 and precision == 1) else None
    masks = np.zeros((n_spectra, (n + 31) // 32), np.uint32)
    counts = np.zeros(n_spectra, np.uint32)
    lib().orc_pipeline(n, sample_rate, enob, kind, 1 if correct_dc else 0, K, threshold, use_window_bins,
                       dc_ignore_window, db_variant, _p(window), _p(raw), n_spectra, precision,
                       _p(spectra), _p(spectra64), _p(masks), _p(counts), threads)
    return {"spectra_db": spectra, "spectra_db64": spectra64, "hit_mask": masks, "hit_count": counts}


def time_domain(raw: np.ndarray, n: int, enob: int, kind: int, correct_dc: bool, threshold: float):
    raw = np.ascontiguousarray(raw)
    nb = raw.nbytes // (n * bytes_per_sample(kind))
    trig = np.zeros(nb, np.uint32)
    mm = np.zeros((nb, 2), np.float32)
    lib().orc_time_domain(n, enob, kind, 1 if correct_dc else 0, threshold, _p(raw), nb, _p(trig), _p(mm))
    return trig, mm


def bench(raw: np.ndarray, n: int, sample_rate: int, enob: int, kind: int, correct_dc: bool, averaging: int,
          threshold: float, window: np.ndarray, use_window_bins: int, dc_ignore_window: int = 4,
          repeats: int = 1, threads: int = 0, faithful=True):
    """Returns (seconds, total_hits, threads_used).  faithful: False = fused CPU path, True = the reference's
    per-buffer copies kept, 2 = the same with the FFTs run eight at a time per worker (SIMD across buffers)."""
    raw = np.ascontiguousarray(raw)
    nb = raw.nbytes // (n * bytes_per_sample(kind))
    window = np.ascontiguousarray(window, np.float32)
    hits = C.c_uint64(0)
    if threads == 0:
        threads = int(lib().orc_hardware_threads()) or 1
    sec = lib().orc_bench(n, sample_rate, enob, kind, 1 if correct_dc else 0, max(1, averaging), threshold,
                          use_window_bins, dc_ignore_window, _p(window), _p(raw), nb, repeats, threads,
                          int(faithful), C.byref(hits))
    return float(sec), int(hits.value), threads


def hackrf_prepass(transfers: np.ndarray, valid_length: int):
    """Returns (patched copy, frequency_hz uint64[T], status uint32[T]) -- hackRFSource.cpp:186-222."""
    buf = np.array(transfers, dtype=np.uint8, copy=True).reshape(-1)
    nt = buf.size // valid_length
    freq = np.zeros(nt, np.uint64)
    status = np.zeros(nt, np.uint32)
    lib().orc_hackrf_prepass(_p(buf), nt, valid_length, _p(freq), _p(status))
    return buf.reshape(nt, valid_length), freq, status


def hardware_threads() -> int:
    return int(lib().orc_hardware_threads()) or 1
