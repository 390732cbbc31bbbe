"""scanner_b200 -- B200-native (sm_100a) spectrum-sense hot path of wpats/scanner.

The product is the C-ABI shared library ``libscanner_b200.so`` (include/scanner_b200.h)
plus the C++ plugin surface under ``csrc/host``.  This Python package is only the ctypes
binding the tests and ``bench.py`` drive it through; it holds no arithmetic of its own and
there is no CPU fallback: every call fails loudly when the CUDA library is missing.
"""
from .binding import (  # noqa: F401
    KIND_BYTE_COMPLEX, KIND_SHORT, KIND_SHORT_COMPLEX, KIND_FLOAT_COMPLEX,
    WIN_HAMMING, WIN_HANN, WIN_BLACKMAN, WIN_RECTANGULAR, WIN_BLACKMAN_HARRIS,
    MODE_TIME_DOMAIN, MODE_FREQUENCY_DOMAIN, OUT_SPECTRUM, OUT_HITS,
    ScannerError, SpectrumSense, lib, lib_path, hit_dtype,
    use_window, hit_frequency, frequency_table, window_build, shard_steps,
    bytes_per_sample, RecordExchange,
)
from .sweep import ShardPlan, plan_shard, gather_step_records, open_record_exchange  # noqa: F401,E402
