"""GPU, BASELINE.json-sized batches: size-independent properties of the device-resident path
(scn_launch_device), where the oracle would take minutes.

  * Parseval: sum_k |X[k]|^2 == N * sum_n |w[n] (x[n] - dc)|^2  -- ties the dB spectrum of EVERY buffer
    to a plain elementwise reduction of the raw samples (no FFT in the checker);
  * permutation equivariance: processing the buffers in a shuffled order permutes the outputs
    (exercises the persistent tile scheduler / prefetch pipeline over thousands of CTAs);
  * determinism, count == popcount(mask), mask inside the candidate band, records sorted;
  * a random sample of spectra against the oracle (bit-exact masks).
"""
import numpy as np
import pytest

import oracle as O
import scanner_b200 as S
from tests import synth

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def device_run(ss, raw_t, n_spectra, n, hit_cap):
    dev = raw_t.device
    spec = torch.empty((n_spectra, n), dtype=torch.float32, device=dev)
    mask = torch.empty((n_spectra, n // 32), dtype=torch.int32, device=dev)
    cnt = torch.empty((n_spectra,), dtype=torch.int32, device=dev)
    hits = torch.zeros((n_spectra, hit_cap, 2), dtype=torch.int32, device=dev)
    ss.launch_device(raw_t.data_ptr(), n_spectra, spec.data_ptr(), mask.data_ptr(), cnt.data_ptr(), hits.data_ptr(),
                     0, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return spec, mask, cnt, hits


@pytest.mark.parametrize("name,kind,enob,dc,n,K,n_spectra", [
    ("cfg2 int8 2048 x 50 steps x 400", 1, 8, True, 2048, 1, 50 * 400),
    ("cfg3 int16 4096 K=64 x 24 steps x 6", 3, 12, False, 4096, 64, 24 * 6),
    ("cfg4 fp32 8192 x 133 steps x 8", 4, 0, False, 8192, 1, 133 * 8),
    ("cfg1 int16 1024 K=16 x 2000", 3, 12, False, 1024, 16, 2000),
    # cluster kernel (scn_cluster.cu): several transforms per persistent cluster -- mbarrier phases flip, the next
    # buffer is prefetched behind the current one, DSMEM regions are reused (148 / 74 / 37 clusters on a B200)
    ("cfg5 fp32 2^16 x 100", 4, 0, False, 1 << 16, 1, 100),
    ("cfg5 int8 2^15 DC x 230", 1, 8, True, 1 << 15, 1, 230),
    ("cfg5 int16 2^14 K=3 x 330", 3, 12, True, 1 << 14, 3, 330),
    ("cfg5 fp32 2^15 K=2 x 160", 4, 0, False, 1 << 15, 2, 160),
])
def test_fullsize_properties(name, kind, enob, dc, n, K, n_spectra):
    dev = torch.device("cuda", 0)
    # a pool of distinct synthetic buffers, tiled to the full batch with a random map (keeps host generation cheap)
    pool = synth.make_buffers(kind, n, (64 if n <= 8192 else 12) * K if K > 1 else (256 if n <= 8192 else 32), enob, seed=31 + n)
    rng = np.random.default_rng(7)
    pool_t = torch.from_numpy(pool).to(dev)
    per = pool.shape[0] // K                       # distinct spectra in the pool
    pick = rng.integers(0, per, n_spectra)
    idx = (pick[:, None] * K + np.arange(K)[None, :]).reshape(-1)
    raw_t = pool_t[torch.from_numpy(idx).to(dev)].contiguous()
    window = S.window_build(S.WIN_HANN if kind == 4 else S.WIN_BLACKMAN_HARRIS, n)
    use_w = S.use_window(0.75, n)
    truth = O.pipeline(pool, n, 8_000_000, enob, kind, dc, K, 0.0, window, use_w, precision=1, want_f64=True)
    thr = synth.guard_banded_threshold(truth["spectra_db64"], n, use_w, quantile=0.995)
    truth = O.pipeline(pool, n, 8_000_000, enob, kind, dc, K, thr, window, use_w, precision=1, want_f64=True)
    cap = 32
    with S.SpectrumSense(n, 8_000_000, enob, thr, window, sample_kind=kind, correct_dc_offset=dc, averaging=K,
                         max_spectra=16, max_hits_per_spectrum=cap) as ss:
        spec, mask, cnt, hits = device_run(ss, raw_t, n_spectra, n, cap)
        spec2, mask2, cnt2, _ = device_run(ss, raw_t, n_spectra, n, cap)
        perm = torch.from_numpy(rng.permutation(n_spectra)).to(dev)
        pidx = (perm[:, None] * K + torch.arange(K, device=dev)[None, :]).reshape(-1)
        spec_p, mask_p, cnt_p, _ = device_run(ss, raw_t[pidx].contiguous(), n_spectra, n, cap)
    # determinism + permutation equivariance (bitwise)
    assert torch.equal(spec.view(torch.int32), spec2.view(torch.int32)) and torch.equal(mask, mask2)
    assert torch.equal(spec[perm].view(torch.int32), spec_p.view(torch.int32))
    assert torch.equal(mask[perm], mask_p) and torch.equal(cnt[perm], cnt_p)
    # every spectrum equals the oracle's result for its pool entry: masks bit-exact, dB within 1e-3 on strong bins
    pick_t = torch.from_numpy(pick).to(dev)
    want_mask = torch.from_numpy(truth["hit_mask"].view(np.int32)).to(dev)[pick_t]
    assert torch.equal(mask, want_mask)
    want_cnt = torch.from_numpy(truth["hit_count"].astype(np.int32)).to(dev)[pick_t]
    assert torch.equal(cnt, want_cnt) and int(cnt.sum()) > 0
    t64 = torch.from_numpy(truth["spectra_db64"]).to(dev)
    mag = 10.0 ** (t64 / 10.0)
    floor = 10.0 * torch.log10(torch.sqrt((mag ** 2).mean(dim=1, keepdim=True))) - 10.0
    strong = (t64 >= floor)[pick_t]
    err = (spec.double() - t64[pick_t]).abs()
    assert float(err[strong].max()) < (1e-3 if n <= 8192 else 2e-3)     # one more twiddle stage above 2^13
    # Parseval per spectrum, straight from the raw samples (no FFT in the checker)
    w_t = torch.from_numpy(window.astype(np.float64)).to(dev)
    x = raw_t.double()
    if kind == 2:
        x = x.permute(0, 2, 1)
    if kind != 4:
        bits_ = 8 if kind == 1 else 16                      # intK_t(1 << (enob-1)) wraps (utility.cpp:40,64)
        mx = 1 << (enob - 1)
        mx = mx - (1 << bits_) if mx >= (1 << (bits_ - 1)) else mx
        scale = 1.0 / float(mx)
        if dc:
            sums = raw_t.to(torch.int64).sum(dim=1)                                   # [B, 2]
            dcv = ((sums & 0xFFFFFFFF) // n).to(torch.int64)                          # uint32(sum) / N
            dcv = torch.where(dcv >= 2 ** 31, dcv - 2 ** 32, dcv).double()
            x = x - dcv[:, None, :]
        x = x * scale
    e_time = ((x * w_t[None, :, None]) ** 2).sum(dim=(1, 2)).reshape(n_spectra, K).mean(dim=1) * n
    e_freq = (10.0 ** (spec.double() / 5.0)).sum(dim=1)                               # |X|^2 = 10^(dB/5)
    assert float(((e_freq - e_time).abs() / e_time).max()) < 2e-5
    # counts / records
    bits = torch.from_numpy(np.unpackbits(mask.cpu().numpy().view(np.uint8), axis=1, bitorder="little"))
    assert torch.equal(bits.sum(dim=1).to(torch.int32), cnt.cpu())
    cand = np.zeros(n, bool)
    cand[(synth.candidate_bins(n, use_w) + n // 2) % n] = True                        # shifted indices
    assert not bits[:, ~cand].any()
    h = hits.cpu().numpy()
    for s in rng.integers(0, n_spectra, 200):
        c = min(int(cnt[s]), cap)
        b = h[s, :c, 0].astype(np.int64)
        assert np.all(np.diff(b) > 0) and np.array_equal(b, np.flatnonzero(bits[s].numpy())[:c])
