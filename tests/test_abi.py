"""CPU: the C-ABI library loads, exports every symbol include/scanner_b200.h declares, its host
helpers restate the reference arithmetic, and it FAILS LOUDLY without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

import oracle as O
import scanner_b200 as S
from scanner_b200 import binding as B
from tests import golden_util as GU
from tests.conftest import HAS_GPU, ROOT

G = GU.load()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "scanner_b200.h")).read()
    return sorted(set(re.findall(r"SCN_API\s+[\w\s\*]+?\b(scn_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = declared_symbols()
    assert len(names) >= 25
    handle = ctypes.CDLL(S.lib_path())
    for name in names:
        assert hasattr(handle, name), f"{name} declared in include/scanner_b200.h but not exported"
    assert set(names) == {s[0] for s in B.SYMBOLS}, "binding.py and the header disagree"
    assert b"sm_100a" in S.lib().scn_version()


def test_product_path_never_imports_the_oracle():
    # the package, the headers and the developer tools: only tests/, bench.py's cpu_baseline / reference legs and
    # __graft_entry__.smoke() may touch oracle/
    for top in ("scanner_b200", "include", "tools"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".sh")):
                    text = open(os.path.join(dirpath, f), errors="ignore").read()
                    assert "import oracle" not in text and "liboracle" not in text and "scanner_oracle" not in text, f


@pytest.mark.skipif(HAS_GPU, reason="checks the no-GPU failure mode")
def test_no_gpu_fails_loudly():
    w = S.window_build(S.WIN_HANN, 1024)
    with pytest.raises(S.ScannerError) as e:
        S.SpectrumSense(1024, 8_000_000, 12, 10.0, w)
    assert e.value.status in (2, 3)       # SCN_ERR_NO_DEVICE / SCN_ERR_CUDA -- never a silent CPU path


def test_host_helpers_match_reference_vectors():
    for i, c in enumerate(G["ft_cases"]):
        np.testing.assert_array_equal(S.frequency_table(int(c[0]), c[1], c[2], c[3], c[4]), G[f"ft_{i}"])
    for n in (256, 1000, 1024, 2048, 8192, 65536):
        assert S.use_window(0.75, n) == O.use_window(0.75, n) == int(0.75 * n / 2.0)
        for wt in (S.WIN_HAMMING, S.WIN_HANN, S.WIN_BLACKMAN, S.WIN_RECTANGULAR, S.WIN_BLACKMAN_HARRIS):
            np.testing.assert_array_equal(S.window_build(wt, n), O.window_build(wt, n))
    rng = np.random.default_rng(1)
    for _ in range(200):
        fs = int(rng.choice([8_000_000, 20_000_000, 56_000_000, 10_000_000, 2_400_000]))
        n = int(rng.choice([256, 1024, 2048, 4096, 8192]))
        c = float(rng.uniform(50e6, 6e9))
        i = int(rng.integers(0, n))
        assert S.hit_frequency(c, fs, n, i) == O.hit_frequency(c, fs, n, i)
    # the reference's own printed frequencies decode back through scn_hit_frequency
    case = next(c for c in GU.scan_cases(G) if c["name"] == "i8_dc_2048")
    f0 = GU.parse_hits(case["text"])[0][0]
    lo, _ = GU.accepted_range(case)
    assert any(S.hit_frequency(case["freqs"][lo], case["fs"], case["n"], i) == f0 for i in range(case["n"]))


def test_window_is_symmetric_and_hann_endpoints():
    w = S.window_build(S.WIN_HANN, 1024)
    assert w[0] == 0.0 and abs(w[-1]) < 1e-7 and np.allclose(w, w[::-1], atol=1e-7)
    bh = S.window_build(S.WIN_BLACKMAN_HARRIS, 2048)
    assert abs(bh[0] - 6e-5) < 1e-6 and abs(bh.max() - 1.0) < 1e-5


def test_shard_steps_partitions():
    for n_steps in (1, 7, 50, 133, 1000):
        for world in (1, 2, 4, 8):
            spans = [S.shard_steps(n_steps, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n_steps
            assert all(spans[r][1] == spans[r + 1][0] for r in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    assert [S.shard_steps(133, r, 8) for r in range(8)][0] == (0, 16)


def test_frequency_table_random_sweeps_agree_with_oracle_and_reference_tool():
    """frequencyTable.cpp:9-37 over random sweeps: C ABI helper == oracle, and == the reference's own code when
    oracle/_ref is present (the build container)."""
    from oracle import ref as R
    rng = np.random.default_rng(7)
    for k in range(60):
        fs = int(rng.choice([2_400_000, 8_000_000, 10_000_000, 20_000_000, 56_000_000]))
        start = float(rng.integers(50, 5000)) * 1e6
        span = float(rng.integers(1, 1200)) * 1e6
        stop = 0.0 if k % 10 == 0 else start + span
        use_bw = float(rng.choice([0.75, 0.5, 0.9]))
        dc_ignore = float(rng.choice([0.0, 0.0, 0.05]))
        a = S.frequency_table(fs, start, stop, use_bw, dc_ignore)
        b = O.frequency_table(fs, start, stop, use_bw, dc_ignore)
        np.testing.assert_array_equal(a, b)
        assert len(a) >= 1
        if R.available() and k < 12:
            np.testing.assert_array_equal(a, R.frequency_table(fs, start, stop, use_bw, dc_ignore))
