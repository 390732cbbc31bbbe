"""Small invocations of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scanner_b200 as S
from tests import synth
cases = [(1, 2048, 8, True, 1), (1, 2048, 8, False, 1),          # WPT
         (1, 8192, 8, True, 1), (3, 4096, 12, True, 3), (4, 8192, 0, False, 1), (4, 4096, 0, False, 2),   # P64
         (3, 1024, 12, True, 2), (2, 512, 12, True, 1), (1, 256, 8, True, 1), (4, 16384, 0, False, 1),    # generic
         (1, 32768, 8, True, 1)]                                                                              # four-step
for kind, n, enob, dc, K in cases:
    raw = synth.make_buffers(kind, n, 5 * K, enob, seed=n + kind)
    w = S.window_build(5, n)
    with S.SpectrumSense(n, 8_000_000, enob, 8.0, w, sample_kind=kind, correct_dc_offset=dc, averaging=K,
                         max_spectra=5, max_hits_per_spectrum=64) as ss:
        r = ss.process(raw)
        print(ss.kernel_name, int(r["hit_count"].sum()), flush=True)
raw = synth.make_buffers(1, 4096, 6, 8, seed=3)
with S.SpectrumSense(4096, 8_000_000, 8, -5.0, None, sample_kind=1, correct_dc_offset=True, mode=S.MODE_TIME_DOMAIN, max_spectra=6) as ss:
    print("time domain", ss.process(raw)["hit_count"].tolist())
# fp32 K = 1 at N = 4096 (next-buffer prefetch into the dead imaginary registers) and an asymmetric window at N = 8192
for n, skew in [(4096, False), (8192, True)]:
    raw = synth.make_buffers(4, n, 7, 0, seed=n)
    w = S.window_build(5, n)
    if skew:
        w = (w * np.linspace(0.8, 1.2, n)).astype(np.float32)
    with S.SpectrumSense(n, 8_000_000, 0, 8.0, w, sample_kind=4, max_spectra=7, max_hits_per_spectrum=64) as ss:
        print(ss.kernel_name, "skew" if skew else "", int(ss.process(raw)["hit_count"].sum()), flush=True)
# standalone converters (every kind, DC on) and the HackRF sweep-frame pre-pass
import torch
for kind, enob in [(1, 8), (3, 12), (2, 12), (4, 0)]:
    raw = synth.make_buffers(kind, 1024, 9, enob, seed=40 + kind)
    with S.SpectrumSense(1024, 8_000_000, enob, 8.0, S.window_build(5, 1024), sample_kind=kind,
                         correct_dc_offset=kind != 4) as ss:
        print("convert kind", kind, float(np.abs(ss.convert(raw)).max()), flush=True)
stream = np.zeros((5, 65536), np.uint8)
stream[:, 0:2] = 0x7F
stream[:, 2] = np.arange(5)
stream[1, 10:12] = 0x7F
with S.SpectrumSense(1024, 8_000_000, 8, 8.0, S.window_build(5, 1024), sample_kind=1, correct_dc_offset=True) as ss:
    d = torch.from_numpy(stream).cuda()
    f = torch.zeros(5, dtype=torch.int64, device="cuda")
    st = torch.zeros(5, dtype=torch.int32, device="cuda")
    ss.hackrf_prepass_device(d.data_ptr(), 5, 65536, f.data_ptr(), st.data_ptr())
    torch.cuda.synchronize()
    print("hackrf prepass", f.cpu().tolist(), st.cpu().tolist(), flush=True)
