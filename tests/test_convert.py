"""Standalone converters: oracle and GPU against the reference's own
utility.cpp on ABI-sized buffers (tests/golden/convert_vectors.npz, make_golden_convert.py).  Bit exact."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle as O   # noqa: E402

G = np.load(os.path.join(ROOT, "tests", "golden", "convert_vectors.npz"), allow_pickle=False)
CASES = [(str(a), int(b), int(c), int(d), bool(int(e))) for a, b, c, d, e in G["cases"]]


@pytest.mark.parametrize("name,kind,n,enob,dc", CASES, ids=[c[0] for c in CASES])
def test_oracle_converters_bit_exact(name, kind, n, enob, dc):
    raw, want = G[f"{name}_raw"], G[f"{name}_out"]
    for b in range(raw.shape[0]):
        got = O.convert(kind, raw[b], n, enob, dc)
        assert np.array_equal(got.view(np.uint32), want[b].view(np.uint32))


def test_quirk_cases_really_take_the_unsigned_path():
    assert np.abs(G["i8_2048_dc_negsum_quirk_out"]).max() > 1000      # 2^32 / N pedestal, utility.cpp:49-50
    assert np.abs(G["i16_256_dc_negsum_quirk_out"]).max() > 1000
    assert np.abs(G["i8_256_dc_out"]).max() <= 2.0


@pytest.mark.gpu
@pytest.mark.parametrize("name,kind,n,enob,dc", CASES, ids=[c[0] for c in CASES])
def test_gpu_converters_bit_exact(name, kind, n, enob, dc):
    import torch
    import scanner_b200 as S
    raw, want = G[f"{name}_raw"], G[f"{name}_out"]
    nb = raw.shape[0]
    with S.SpectrumSense(n, 8_000_000, enob, 0.0, S.window_build(5, n), sample_kind=kind, correct_dc_offset=dc,
                         max_spectra=4) as ss:
        got = ss.convert(raw)                                         # scn_convert_host
        assert got.shape == (nb, n)
        assert np.array_equal(got.view(np.float32).reshape(nb, n, 2).view(np.uint32), want.view(np.uint32))
        d_raw = torch.from_numpy(np.ascontiguousarray(raw).view(np.uint8).reshape(-1).copy()).cuda()
        d_out = torch.zeros((nb, n, 2), dtype=torch.float32, device="cuda")
        ss.convert_device(d_raw.data_ptr(), nb, d_out.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert np.array_equal(d_out.cpu().numpy().view(np.uint32), want.view(np.uint32))


@pytest.mark.gpu
def test_gpu_float_complex_passes_through():
    import scanner_b200 as S
    rng = np.random.default_rng(5)
    raw = rng.standard_normal((3, 512, 2)).astype(np.float32)
    with S.SpectrumSense(512, 8_000_000, 0, 0.0, S.window_build(5, 512), sample_kind=S.KIND_FLOAT_COMPLEX) as ss:
        got = ss.convert(raw)
    assert np.array_equal(got.view(np.float32).reshape(3, 512, 2), raw)      # messageQueue.h:231-237


@pytest.mark.gpu
def test_gpu_convert_many_buffers_matches_oracle():
    """More buffers than CTAs in flight (grid-stride path), every kind."""
    import scanner_b200 as S
    from tests import synth
    for kind, enob, dc in [(1, 8, True), (3, 12, True), (2, 12, True)]:
        n = 2048
        raw = synth.make_buffers(kind, n, 1500, enob, seed=900 + kind)
        with S.SpectrumSense(n, 8_000_000, enob, 0.0, S.window_build(5, n), sample_kind=kind,
                             correct_dc_offset=dc) as ss:
            got = ss.convert(raw).view(np.float32).reshape(1500, n, 2)
        for b in (0, 1, 777, 1499):
            assert np.array_equal(got[b].view(np.uint32), O.convert(kind, raw[b], n, enob, dc).view(np.uint32))
