"""ctypes binding of include/scanner_b200.h (no arithmetic lives here)."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

KIND_BYTE_COMPLEX, KIND_SHORT, KIND_SHORT_COMPLEX, KIND_FLOAT_COMPLEX = 1, 2, 3, 4
WIN_HAMMING, WIN_HANN, WIN_BLACKMAN, WIN_RECTANGULAR, WIN_BLACKMAN_HARRIS = 0, 1, 2, 3, 5
MODE_TIME_DOMAIN, MODE_FREQUENCY_DOMAIN = 1, 2
OUT_SPECTRUM, OUT_HITS = 1, 2

_STATUS = {1: "SCN_ERR_INVALID", 2: "SCN_ERR_NO_DEVICE", 3: "SCN_ERR_CUDA", 4: "SCN_ERR_CAPACITY",
           5: "SCN_ERR_BUSY", 6: "SCN_ERR_ALIGNMENT"}

hit_dtype = np.dtype([("bin", np.uint32), ("power_db", np.float32)])


def bytes_per_sample(kind: int) -> int:
    return {KIND_BYTE_COMPLEX: 2, KIND_SHORT: 4, KIND_SHORT_COMPLEX: 4, KIND_FLOAT_COMPLEX: 8}[kind]


class ScannerError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"{_STATUS.get(status, status)}: {message}")
        self.status = status


class _Config(C.Structure):
    _fields_ = [
        ("device", C.c_int32), ("sample_count", C.c_uint32), ("sample_rate", C.c_uint32),
        ("enob", C.c_uint32), ("sample_kind", C.c_uint32), ("correct_dc_offset", C.c_uint32),
        ("averaging", C.c_uint32), ("mode", C.c_uint32), ("threshold", C.c_float),
        ("use_window", C.c_uint32), ("dc_ignore_window", C.c_uint32),
        ("window", C.POINTER(C.c_float)), ("max_spectra", C.c_uint32),
        ("max_hits_per_spectrum", C.c_uint32), ("flags", C.c_uint32), ("ticket_slots", C.c_uint32),
    ]


def lib_path() -> str:
    # SCN_LIB: developer knob to time an experiment build (tools/build_variant.sh); never a fallback
    override = os.environ.get("SCN_LIB")
    if override:
        return override
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "libscanner_b200.so")


_lib = None

# every symbol include/scanner_b200.h declares: (name, restype, argtypes)
_VP, _U32, _U64, _F, _D, _I = C.c_void_p, C.c_uint32, C.c_uint64, C.c_float, C.c_double, C.c_int
SYMBOLS = [
    ("scn_version", C.c_char_p, []),
    ("scn_device_count", _I, [C.POINTER(_I)]),
    ("scn_last_error", C.c_char_p, []),
    ("scn_create", _I, [C.POINTER(_Config), C.POINTER(_VP)]),
    ("scn_destroy", _I, [_VP]),
    ("scn_set_threshold", _I, [_VP, _F]),
    ("scn_buffer_bytes", C.c_size_t, [_VP]),
    ("scn_mask_words", _U32, [_VP]),
    ("scn_alloc_pinned", _I, [C.c_size_t, C.POINTER(_VP)]),
    ("scn_free_pinned", _I, [_VP]),
    ("scn_process_host", _I, [_VP, _VP, _U32, _VP, _VP, _VP, _VP, _VP]),
    ("scn_submit", _I, [_VP, _VP, _U32, C.POINTER(_U32)]),
    ("scn_submit_gather", _I, [_VP, C.POINTER(_VP), C.POINTER(_U32), _U32, _U32, C.POINTER(_U32)]),
    ("scn_collect", _I, [_VP, _U32, _VP, _VP, _VP, _VP, _VP]),
    ("scn_collect_view", _I, [_VP, _U32, C.POINTER(_VP), C.POINTER(_VP), C.POINTER(_VP), C.POINTER(_VP)]),
    ("scn_launch_device", _I, [_VP, _VP, _U32, _VP, _VP, _VP, _VP, _VP, _VP]),
    ("scn_launch_count", _U64, [_VP]),
    ("scn_kernel_name", C.c_char_p, [_VP]),
    ("scn_kernel_info", _I, [_VP] + [C.POINTER(_I)] * 5),
    ("scn_record_words", _U32, [_VP]),
    ("scn_summarize_steps", _I, [_VP, _VP, _VP, _U32, _U64, _U32, _U32, _VP, _VP]),
    ("scn_merge_step_records", _I, [_VP, _VP, _U32, _U32, _VP, _VP]),
    ("scn_exchange_create", _I, [_I, _U32, _U32, _U32, _U32, C.POINTER(_VP)]),
    ("scn_exchange_handle", _I, [_VP, _VP]),
    ("scn_exchange_connect_ipc", _I, [_VP, _VP]),
    ("scn_exchange_connect_local", _I, [C.POINTER(_VP), _U32]),
    ("scn_exchange_publish", _I, [_VP, _VP, _VP, C.POINTER(_U64)]),
    ("scn_exchange_merge", _I, [_VP, _U64, _VP, _VP]),
    ("scn_exchange_step", _I, [_VP, _VP, _VP, _VP, C.POINTER(_U64)]),
    ("scn_exchange_publish_host", _I, [_VP, _VP, C.POINTER(_U64)]),
    ("scn_exchange_merge_host", _I, [_VP, _U64, _VP]),
    ("scn_nccl_gather_create", _I, [C.POINTER(_I), _U32, _U32, _U32, C.POINTER(_VP)]),
    ("scn_nccl_gather_merge_host", _I, [_VP, C.POINTER(_VP), _VP]),
    ("scn_nccl_gather_destroy", _I, [_VP]),
    ("scn_exchange_status", _I, [_VP, C.POINTER(_U32)]),
    ("scn_exchange_slots", _U32, []),
    ("scn_exchange_destroy", _I, [_VP]),
    ("scn_hackrf_prepass_device", _I, [_VP, _VP, _U32, _U32, _VP, _VP, _VP]),
    ("scn_convert_device", _I, [_VP, _VP, _U32, _VP, _VP]),
    ("scn_convert_host", _I, [_VP, _VP, _U32, _VP]),
    ("scn_use_window", _U32, [_D, _U32]),
    ("scn_hit_frequency", _U64, [_D, _U32, _U32, _U32]),
    ("scn_frequency_table", _U32, [_U32, _D, _D, _D, _D, C.POINTER(_D), _U32]),
    ("scn_window_build", _I, [_I, _U32, C.POINTER(_F)]),
    ("scn_shard_steps", None, [_U32, _U32, _U32, C.POINTER(_U32), C.POINTER(_U32)]),
]


def lib() -> C.CDLL:
    """Loads libscanner_b200.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        path = lib_path()
        if not os.path.exists(path):
            raise ScannerError(3, f"{path} is missing: run `make` (or __graft_entry__.build()); "
                                  "there is no CPU fallback")
        handle = C.CDLL(path)
        for name, restype, argtypes in SYMBOLS:
            if os.environ.get("SCN_LIB") and not hasattr(handle, name):
                continue      # an experiment build of an older revision (A/B timing only): newer entry points are absent
            fn = getattr(handle, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = handle
    return _lib


def _check(status: int) -> None:
    if status != 0:
        raise ScannerError(status, lib().scn_last_error().decode())


# ---- host helpers (pure integer/double arithmetic restated from the reference) -------------

def use_window(use_bandwidth: float, sample_count: int) -> int:
    return int(lib().scn_use_window(use_bandwidth, sample_count))


def hit_frequency(center: float, sample_rate: int, sample_count: int, bin_index: int) -> int:
    return int(lib().scn_hit_frequency(center, sample_rate, sample_count, bin_index))


def frequency_table(sample_rate: int, start: float, stop: float, use_bandwidth: float = 0.75,
                    dc_ignore_width: float = 0.0) -> np.ndarray:
    n = lib().scn_frequency_table(sample_rate, start, stop, use_bandwidth, dc_ignore_width, None, 0)
    out = np.zeros(n, dtype=np.float64)
    lib().scn_frequency_table(sample_rate, start, stop, use_bandwidth, dc_ignore_width,
                              out.ctypes.data_as(C.POINTER(_D)), n)
    return out


def window_build(win_type: int, sample_count: int) -> np.ndarray:
    out = np.zeros(sample_count, dtype=np.float32)
    _check(lib().scn_window_build(win_type, sample_count, out.ctypes.data_as(C.POINTER(_F))))
    return out


def shard_steps(n_steps: int, rank: int, world: int) -> tuple[int, int]:
    b, e = _U32(0), _U32(0)
    lib().scn_shard_steps(n_steps, rank, world, C.byref(b), C.byref(e))
    return int(b.value), int(e.value)


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(_VP)


class SpectrumSense:
    """One scn_ctx.  Mirrors ProcessSamples' constructor arguments (process.h:74-85) plus the
    SampleQueue's (kind, enob, correctDCOffset) (messageQueue.h:141-146)."""

    def __init__(self, sample_count: int, sample_rate: int, enob: int, threshold: float,
                 window: Optional[np.ndarray], sample_kind: int = KIND_SHORT_COMPLEX,
                 correct_dc_offset: bool = False, averaging: int = 1,
                 mode: int = MODE_FREQUENCY_DOMAIN, use_bandwidth: float = 0.75,
                 dc_ignore_window: int = 4, max_spectra: int = 1024,
                 max_hits_per_spectrum: int = 0, flags: int = OUT_SPECTRUM | OUT_HITS,
                 device: int = 0, ticket_slots: int = 2, use_window_bins: Optional[int] = None):
        self._lib = lib()
        self.N = int(sample_count)
        self.kind = int(sample_kind)
        self.K = max(1, int(averaging))
        self.mode = mode
        self.flags = flags
        self.sample_rate = int(sample_rate)
        self._window = None if window is None else np.ascontiguousarray(window, dtype=np.float32)
        if self._window is not None and self._window.shape != (self.N,):
            raise ScannerError(1, f"window must have {self.N} taps")
        cfg = _Config()
        cfg.device = device
        cfg.sample_count = self.N
        cfg.sample_rate = self.sample_rate
        cfg.enob = enob
        cfg.sample_kind = self.kind
        cfg.correct_dc_offset = 1 if correct_dc_offset else 0
        cfg.averaging = self.K
        cfg.mode = mode
        cfg.threshold = threshold
        cfg.use_window = use_window(use_bandwidth, self.N) if use_window_bins is None else use_window_bins
        cfg.dc_ignore_window = dc_ignore_window
        cfg.window = None if self._window is None else self._window.ctypes.data_as(C.POINTER(C.c_float))
        cfg.max_spectra = max_spectra
        cfg.max_hits_per_spectrum = max_hits_per_spectrum
        cfg.flags = flags
        cfg.ticket_slots = ticket_slots
        self.use_window = int(cfg.use_window)
        self.max_spectra = max_spectra
        self.hit_cap = (max_hits_per_spectrum or self.N) if (flags & OUT_HITS) else 0
        handle = _VP()
        _check(self._lib.scn_create(C.byref(cfg), C.byref(handle)))
        self._ctx = handle
        self.words = self.N // 32
        self.buffer_bytes = int(self._lib.scn_buffer_bytes(self._ctx))

    # -- lifetime -------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_ctx", None):
            self._lib.scn_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- facts ----------------------------------------------------------------------------
    @property
    def kernel_name(self) -> str:
        return self._lib.scn_kernel_name(self._ctx).decode()

    @property
    def launch_count(self) -> int:
        return int(self._lib.scn_launch_count(self._ctx))

    def kernel_info(self) -> dict:
        v = [_I(0) for _ in range(5)]
        _check(self._lib.scn_kernel_info(self._ctx, *[C.byref(x) for x in v]))
        return dict(zip(("ctas_per_sm", "threads", "smem_bytes", "regs_per_thread", "grid"),
                        (int(x.value) for x in v)))

    def set_threshold(self, threshold: float) -> None:
        _check(self._lib.scn_set_threshold(self._ctx, threshold))

    # -- host path ------------------------------------------------------------------------
    def _alloc_outputs(self, n: int, want_spectrum: bool, want_hits: bool):
        td = self.mode == MODE_TIME_DOMAIN
        spectra = np.empty((n, self.N), np.float32) if (want_spectrum and not td) else None
        masks = None if td else np.empty((n, self.words), np.uint32)
        counts = np.empty(n, np.uint32)
        hits = np.zeros((n, self.hit_cap), hit_dtype) if (want_hits and self.hit_cap and not td) else None
        tdmm = np.empty((n, 2), np.float32) if td else None
        return spectra, masks, counts, hits, tdmm

    def process(self, raw: np.ndarray, n_spectra: Optional[int] = None, want_spectrum: bool = True,
                want_hits: bool = True) -> dict:
        """Synchronous host call (scn_process_host).  raw: any contiguous array holding
        n_spectra * K buffers."""
        raw = np.ascontiguousarray(raw)
        if n_spectra is None:
            n_spectra = raw.nbytes // (self.buffer_bytes * self.K)
        if raw.nbytes < n_spectra * self.K * self.buffer_bytes:
            raise ScannerError(1, "raw array shorter than n_spectra * K buffers")
        spectra, masks, counts, hits, tdmm = self._alloc_outputs(
            n_spectra, want_spectrum and bool(self.flags & OUT_SPECTRUM), want_hits)
        _check(self._lib.scn_process_host(self._ctx, _ptr(raw), n_spectra, _ptr(spectra), _ptr(masks),
                                          _ptr(counts), _ptr(hits), _ptr(tdmm)))
        return {"spectra_db": spectra, "hit_mask": masks, "hit_count": counts, "hits": hits,
                "td_max_min": tdmm}

    def submit(self, raw_ptr: int, n_spectra: int) -> int:
        ticket = _U32(0)
        _check(self._lib.scn_submit(self._ctx, _VP(raw_ptr), n_spectra, C.byref(ticket)))
        return int(ticket.value)

    def submit_gather(self, run_ptrs: list[int], run_buffers: list[int], n_spectra: int) -> int:
        ticket = _U32(0)
        ptrs = (_VP * len(run_ptrs))(*run_ptrs)
        cnts = (_U32 * len(run_buffers))(*run_buffers)
        _check(self._lib.scn_submit_gather(self._ctx, ptrs, cnts, len(run_ptrs), n_spectra, C.byref(ticket)))
        return int(ticket.value)

    def collect(self, ticket: int, spectra=None, masks=None, counts=None, hits=None, tdmm=None) -> None:
        _check(self._lib.scn_collect(self._ctx, ticket, _ptr(spectra), _ptr(masks), _ptr(counts),
                                     _ptr(hits), _ptr(tdmm)))

    # -- device path ----------------------------------------------------------------------
    def launch_device(self, d_raw: int, n_spectra: int, d_spectra: int = 0, d_masks: int = 0,
                      d_counts: int = 0, d_hits: int = 0, d_td: int = 0, stream: int = 0) -> None:
        """Enqueue one fused kernel on `stream` over device pointers (ints)."""
        _check(self._lib.scn_launch_device(self._ctx, _VP(d_raw), n_spectra, _VP(d_spectra or None),
                                           _VP(d_masks or None), _VP(d_counts or None),
                                           _VP(d_hits or None), _VP(d_td or None), _VP(stream or None)))


    def summarize_steps(self, d_masks: int, d_counts: int, n_spectra: int, first_unit: int,
                        units_per_step: int, n_steps: int, d_records: int, stream: int = 0) -> None:
        _check(self._lib.scn_summarize_steps(self._ctx, _VP(d_masks or None), _VP(d_counts or None), n_spectra,
                                             first_unit, units_per_step, n_steps, _VP(d_records),
                                             _VP(stream or None)))

    def merge_step_records(self, d_parts: int, n_parts: int, n_steps: int, d_out: int, stream: int = 0) -> None:
        _check(self._lib.scn_merge_step_records(self._ctx, _VP(d_parts), n_parts, n_steps, _VP(d_out),
                                                _VP(stream or None)))

    def convert(self, raw: np.ndarray) -> np.ndarray:
        """raw buffers -> complex64 [n_buffers][N], the reference's converters (utility.cpp:9-84) on the GPU."""
        raw = np.ascontiguousarray(raw)
        nb = raw.nbytes // self.buffer_bytes
        out = np.empty((nb, self.N), np.complex64)
        _check(self._lib.scn_convert_host(self._ctx, _ptr(raw), nb, _ptr(out)))
        return out

    def convert_device(self, d_raw: int, n_buffers: int, d_out: int, stream: int = 0) -> None:
        _check(self._lib.scn_convert_device(self._ctx, _VP(d_raw), n_buffers, _VP(d_out), _VP(stream or None)))

    def hackrf_prepass_device(self, d_transfers: int, n_transfers: int, valid_length: int,
                              d_frequency_hz: int = 0, d_status: int = 0, stream: int = 0) -> None:
        """In-place HackRF sweep-frame pre-pass over device-resident transfers (hackRFSource.cpp:186-222)."""
        _check(self._lib.scn_hackrf_prepass_device(self._ctx, _VP(d_transfers or None), n_transfers, valid_length,
                                                   _VP(d_frequency_hz or None), _VP(d_status or None),
                                                   _VP(stream or None)))

    @property
    def record_words(self) -> int:
        return self.words + 2


IPC_HANDLE_BYTES = 64


class RecordExchange:
    """One scn_exchange: this rank's window for the NVLink peer-memory exchange of per-step records."""

    def __init__(self, device: int, rank: int, world: int, n_steps: int, record_words: int):
        self._lib = lib()
        self.rank, self.world = rank, world
        h = _VP()
        _check(self._lib.scn_exchange_create(device, rank, world, n_steps, record_words, C.byref(h)))
        self._x = h

    def handle(self) -> bytes:
        buf = (C.c_uint8 * IPC_HANDLE_BYTES)()
        _check(self._lib.scn_exchange_handle(self._x, C.cast(buf, _VP)))
        return bytes(buf)

    def connect_ipc(self, handles: list[bytes]) -> None:
        blob = b"".join(handles)
        assert len(blob) == IPC_HANDLE_BYTES * self.world
        buf = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
        _check(self._lib.scn_exchange_connect_ipc(self._x, C.cast(buf, _VP)))

    @staticmethod
    def connect_local(exchanges: "list[RecordExchange]") -> None:
        arr = (_VP * len(exchanges))(*[e._x for e in exchanges])
        _check(lib().scn_exchange_connect_local(arr, len(exchanges)))

    def publish(self, d_records: int, stream: int = 0) -> int:
        seq = _U64(0)
        _check(self._lib.scn_exchange_publish(self._x, _VP(d_records), _VP(stream or None), C.byref(seq)))
        return int(seq.value)

    def step(self, d_records: int, d_merged_previous: int, stream: int = 0) -> int:
        """publish(d_records) + merge(previous batch) in one launch; returns this batch's sequence number."""
        seq = _U64(0)
        _check(self._lib.scn_exchange_step(self._x, _VP(d_records), _VP(d_merged_previous), _VP(stream or None), C.byref(seq)))
        return int(seq.value)

    def merge(self, seq: int, d_merged: int, stream: int = 0) -> None:
        _check(self._lib.scn_exchange_merge(self._x, seq, _VP(d_merged), _VP(stream or None)))

    def status(self) -> int:
        v = _U32(0)
        _check(self._lib.scn_exchange_status(self._x, C.byref(v)))
        return int(v.value)

    def close(self) -> None:
        if getattr(self, "_x", None):
            self._lib.scn_exchange_destroy(self._x)
            self._x = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def alloc_pinned(nbytes: int) -> tuple[int, np.ndarray]:
    """Pinned host bytes as (address, uint8 view)."""
    p = _VP()
    _check(lib().scn_alloc_pinned(nbytes, C.byref(p)))
    buf = (C.c_uint8 * nbytes).from_address(p.value)
    return p.value, np.frombuffer(buf, dtype=np.uint8)


def free_pinned(address: int) -> None:
    _check(lib().scn_free_pinned(_VP(address)))
