"""GPU: N = 2^15 and 2^16 against the oracle, same contract as the in-CTA sizes -- masks / counts / hit order bit-exact,
power within 1e-3 dB -- through BOTH large-N paths: the default single-pass cluster kernel (scn_cluster.cu: one transform
per 2- / 4-CTA cluster in distributed shared memory) and the older four-step path through an HBM intermediate
(scn_large.cu, SCN_FOUR_STEP=1), which also serves as an independent cross-check of the cluster kernel."""
import os
import numpy as np
import pytest

import scanner_b200 as S
from tests.test_gpu_parity import run_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("log2n", [15, 16])
@pytest.mark.parametrize("kind,enob,dc", [(S.KIND_BYTE_COMPLEX, 8, True), (S.KIND_SHORT_COMPLEX, 12, False),
                                          (S.KIND_SHORT, 12, True), (S.KIND_FLOAT_COMPLEX, 0, False)])
def test_large_parity(log2n, kind, enob, dc):
    # one more twiddle stage (W_N in fp32) than the in-CTA path: rms error class 4x the CPU fp32 FFT
    run_case(kind, 1 << log2n, enob, dc, 1, 3, seed=700 + log2n * 10 + kind, acc_factor=4.0)


@pytest.fixture
def four_step(monkeypatch):
    monkeypatch.setenv("SCN_FOUR_STEP", "1")        # read by scn_create


@pytest.mark.parametrize("log2n,kind,enob,dc", [(15, S.KIND_BYTE_COMPLEX, 8, True), (16, S.KIND_FLOAT_COMPLEX, 0, False),
                                                (16, S.KIND_SHORT_COMPLEX, 12, True),
                                                (14, S.KIND_BYTE_COMPLEX, 8, True)])    # 2^14: the 16-point in-CTA kernel
def test_four_step_path_parity(four_step, log2n, kind, enob, dc):
    run_case(kind, 1 << log2n, enob, dc, 1, 3, seed=700 + log2n * 10 + kind, acc_factor=4.0)


def test_four_step_two_tickets_in_flight_do_not_share_scratch(four_step):
    # chunked submits keep two tickets in flight on different streams; the per-context intermediates are serialised
    run_case(S.KIND_SHORT_COMPLEX, 1 << 15, 12, True, 2, 7, seed=811, max_spectra=2, hit_cap=5, acc_factor=4.0)


def test_cluster_kernel_is_the_default_large_path():
    w = S.window_build(S.WIN_HANN, 1 << 16)
    with S.SpectrumSense(1 << 16, 8_000_000, 0, 10.0, w, sample_kind=S.KIND_FLOAT_COMPLEX, max_spectra=1) as ss:
        assert "cluster" in ss.kernel_name and "DSMEM" in ss.kernel_name


@pytest.mark.parametrize("kind,enob,dc,K", [(S.KIND_FLOAT_COMPLEX, 0, False, 1), (S.KIND_SHORT_COMPLEX, 12, True, 3),
                                            (S.KIND_SHORT, 12, False, 1), (S.KIND_BYTE_COMPLEX, 8, True, 1),
                                            (S.KIND_BYTE_COMPLEX, 8, False, 2)])
def test_cluster_kernel_single_cta_size(kind, enob, dc, K):
    run_case(kind, 1 << 14, enob, dc, K, 5, seed=640 + kind, acc_factor=4.0)        # C = 1: 4 rows x 4096 in one CTA


def test_large_averaging_and_chunking():
    # K = 4 averaging; 5 spectra through a context sized for 2 (chunked submits), small record cap
    run_case(S.KIND_SHORT_COMPLEX, 1 << 15, 12, True, 4, 5, seed=801, max_spectra=2, hit_cap=7, acc_factor=4.0)


def test_large_band_edges():
    n = 1 << 16
    rect = S.window_build(S.WIN_RECTANGULAR, n)
    use_w, half = S.use_window(0.75, n), n // 2
    idx = [half - use_w - 1, half - use_w, half + use_w, half + use_w + 1, half - 4, half - 3, half + 3, half + 4, 5, n - 7]
    t = np.arange(n)
    x = np.zeros((len(idx), n, 2), np.float32)
    for s, i in enumerate(idx):
        j = (i + half) % n
        x[s, :, 0] = np.cos(2 * np.pi * ((j * t) % n) / n)
        x[s, :, 1] = np.sin(2 * np.pi * ((j * t) % n) / n)
    with S.SpectrumSense(n, 8_000_000, 0, 20.0, rect, sample_kind=S.KIND_FLOAT_COMPLEX, max_spectra=len(idx)) as ss:
        res = ss.process(x)
    for s, i in enumerate(idx):
        j = (i + half) % n
        want = int((half - use_w) <= i <= (half + use_w) and not (j < 4 or (n - j) < 4))
        assert res["hit_count"][s] == want, (i, res["hit_count"][s])
        if want:
            assert res["hits"]["bin"][s, 0] == i and abs(res["hits"]["power_db"][s, 0] - 10 * np.log10(n)) < 1e-3
            assert res["hit_mask"][s, i >> 5] == np.uint32(1 << (i & 31))


def test_cluster_spectrum_pointer_alignment():
    """The ABI only asks for a 16-byte aligned RAW pointer: a spectra pointer that is merely 4-byte aligned must give
    the same bits and must not write outside its range (guards any future vector / bulk-store epilogue)."""
    torch = pytest.importorskip("torch")
    from tests import synth
    n, ns = 1 << 15, 5
    raw = synth.make_buffers(S.KIND_FLOAT_COMPLEX, n, ns, 0, seed=99)
    dev = torch.device("cuda", 0)
    raw_t = torch.from_numpy(raw).to(dev)
    w = S.window_build(S.WIN_HANN, n)
    with S.SpectrumSense(n, 8_000_000, 0, 12.0, w, sample_kind=S.KIND_FLOAT_COMPLEX, max_spectra=ns) as ss:
        outs = []
        for off in (0, 1, 4):                         # floats: 16-byte aligned, 4-byte aligned, 16-byte aligned again
            buf = torch.zeros(ns * n + 8, dtype=torch.float32, device=dev)
            mask = torch.zeros((ns, n // 32), dtype=torch.int32, device=dev)
            cnt = torch.zeros((ns,), dtype=torch.int32, device=dev)
            ss.launch_device(raw_t.data_ptr(), ns, buf.data_ptr() + 4 * off, mask.data_ptr(), cnt.data_ptr(), 0, 0,
                             torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            outs.append((buf[off:off + ns * n].clone(), mask, cnt))
            assert float(buf[:off].abs().sum()) == 0.0 and float(buf[off + ns * n:].abs().sum()) == 0.0
    for o in outs[1:]:
        assert torch.equal(o[0].view(torch.int32), outs[0][0].view(torch.int32))
        assert torch.equal(o[1], outs[0][1]) and torch.equal(o[2], outs[0][2])
    assert int(outs[0][2].sum()) > 0
