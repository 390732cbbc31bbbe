#!/usr/bin/env python
"""Kernel-only timing of the fused kernel on device-resident random data (developer tool).
usage: kbench.py <kind> <log2n> <dc> <K> [n_buffers] [spectrum 0/1]   (SCN_LIB selects an experiment build)"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scanner_b200 as S
kind, log2n, dc, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
n = 1 << log2n
bps = S.bytes_per_sample(kind)
nbuf = int(sys.argv[5]) if len(sys.argv) > 5 else max(K, ((768 << 20) // (n * bps)) // K * K)
spectrum = int(sys.argv[6]) if len(sys.argv) > 6 else 1
ns = nbuf // K
dev = torch.device("cuda", 0)
if kind == 4:
    raw = (0.05 * torch.randn((nbuf, n, 2), device=dev)).contiguous()
else:
    amp = 20 if kind == 1 else 300
    raw = torch.randint(-amp, amp, (nbuf, n, 2), device=dev, dtype=torch.int8 if kind == 1 else torch.int16)
w = S.window_build(5, n)
ctx = S.SpectrumSense(n, 20_000_000, 8 if kind == 1 else 12, 40.0, w, sample_kind=kind, correct_dc_offset=bool(dc),
                      averaging=K, max_spectra=16, max_hits_per_spectrum=16)
d_spec = torch.empty((ns, n), dtype=torch.float32, device=dev) if spectrum else None
d_mask = torch.empty((ns, n // 32), dtype=torch.int32, device=dev)
d_cnt = torch.empty((ns,), dtype=torch.int32, device=dev)
st = torch.cuda.current_stream()
def go():
    ctx.launch_device(raw.data_ptr(), ns, d_spec.data_ptr() if spectrum else 0, d_mask.data_ptr(), d_cnt.data_ptr(), 0, 0, st.cuda_stream)
for _ in range(3): go()
torch.cuda.synchronize()
ts = []
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st); go(); e1.record(st); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
ms = float(np.median(ts))
samples = nbuf * n
bpsamp = bps + (4.0 / K if spectrum else 0) + (n / 8 + 4) / (K * n)
info = ctx.kernel_info()
print(f"{os.path.basename(os.environ.get('SCN_LIB','default')):28s} kind={kind} N=2^{log2n} dc={dc} K={K} S={spectrum}: {ms:.3f} ms  {samples/ms/1e6:8.1f} Gsamples/s  "
      f"{samples*bpsamp/ms/1e6:7.1f} GB/s ({samples*bpsamp/ms/1e6/6548.2:.3f} of HBM)  regs={info['regs_per_thread']} ctas/SM={info['ctas_per_sm']} hits={int(d_cnt.sum())}", flush=True)
