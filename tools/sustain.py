#!/usr/bin/env python
"""Sustained-clock check (developer tool): loops the fused kernel back to back for <seconds> and samples SM clock /
power through NVML.  usage: sustain.py <kind> <log2n> <dc> <K> [seconds]   (SCN_LIB selects an experiment build)"""
import os, sys, threading, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scanner_b200 as S
kind, log2n, dc, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
seconds = float(sys.argv[5]) if len(sys.argv) > 5 else 4.0
n = 1 << log2n
bps = S.bytes_per_sample(kind)
nbuf = max(K, ((768 << 20) // (n * bps)) // K * K)
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
d_spec = torch.empty((ns, n), dtype=torch.float32, device=dev)
d_mask = torch.empty((ns, n // 32), dtype=torch.int32, device=dev)
d_cnt = torch.empty((ns,), dtype=torch.int32, device=dev)
st = torch.cuda.current_stream()
def go():
    ctx.launch_device(raw.data_ptr(), ns, d_spec.data_ptr(), d_mask.data_ptr(), d_cnt.data_ptr(), 0, 0, st.cuda_stream)
import pynvml
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
clk, pw, halt = [], [], threading.Event()
def sample():
    while not halt.is_set():
        clk.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)); pw.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1e3)
        time.sleep(0.02)
for _ in range(3): go()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
go(); torch.cuda.synchronize()
e0.record(st); go(); e1.record(st); torch.cuda.synchronize()
one = e0.elapsed_time(e1)
reps = max(10, int(seconds * 1e3 / one))
th = threading.Thread(target=sample, daemon=True); th.start()
evs = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
evs[0].record(st)
for i in range(reps):
    go(); evs[i + 1].record(st)
torch.cuda.synchronize()
halt.set(); th.join()
ts = np.array([evs[i].elapsed_time(evs[i + 1]) for i in range(reps)])
samples = nbuf * n
q = len(ts) // 4
print(f"{os.path.basename(os.environ.get('SCN_LIB','default'))} kind={kind} N=2^{log2n}: burst {samples/one/1e6:.1f} Gs/s; sustained {reps} launches "
      f"{ts.sum()/1e3:.2f} s: mean {samples/ts.mean()/1e6:.1f} first-quarter {samples/ts[:q].mean()/1e6:.1f} last-quarter {samples/ts[-q:].mean()/1e6:.1f} Gs/s; "
      f"SM MHz median {np.median(clk):.0f} min {min(clk)} last {clk[-1]}; power W median {np.median(pw):.0f} max {max(pw):.0f}", flush=True)
