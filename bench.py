#!/usr/bin/env python
"""bench.py -- IQ Msamples/s through the fused convert+window+FFT+dB+average+detect path.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W

A "step" is one pass of the hot path over one batch of synthetic IQ: every rank runs the fused
sm_100a kernel over its contiguous shard of the sweep's (retune step, buffer) units, reduces its
detections to per-retune-step records, and (N > 1) all-gathers those small records over NCCL.
Workload at every N: BASELINE.json configs[1] (HackRF-style int8 IQ at 20 MS/s, 2048-pt FFT,
50-step frequencyTable sweep, DC correction on, Blackman-Harris window, K = 1), dB spectrum
written (6 B/sample + mask).  Weak scaling: per-GPU buffers fixed, the
dwell (buffers per retune step) grows with N.

Prints ONE JSON line (rank 0).  `value` times device-resident inputs; `e2e` times the same work
through scn_submit/scn_collect with pinned HOST buffers (H2D and D2H inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "IQ Msamples/s through fused FFT+power+detect"
UNIT = "Msamples/s"

WORKLOADS = {
    # BASELINE.json configs[1]; steps from FrequencyTable(20e6, 2.4e9, 3.15e9) = 50 (frequencyTable.cpp:17-36)
    "cfg2": dict(desc="HackRF-style int8 IQ @20 MS/s, N=2048, 50-step sweep 2.4-3.15 GHz, DC on, "
                      "Blackman-Harris, K=1, dB spectrum + mask + count written",
                 kind=1, enob=8, dc=True, n=2048, K=1, fs=20_000_000, start=2.4e9, stop=3.15e9, win=5,
                 buffers_per_step=4096),
    # BASELINE.json configs[2]
    "cfg3": dict(desc="int16 IQ @56 MS/s, N=4096, K=64 averaging, 24-step sweep 1-2 GHz, DC off",
                 kind=3, enob=12, dc=False, n=4096, K=64, fs=56_000_000, start=1e9, stop=2e9, win=5,
                 buffers_per_step=64 * 48),
    # BASELINE.json configs[3]
    "cfg4": dict(desc="fp32 IQ @10 MS/s, N=8192, Hann, K=1, 133-step sweep 0.1-1.1 GHz",
                 kind=4, enob=0, dc=False, n=8192, K=1, fs=10_000_000, start=0.1e9, stop=1.1e9, win=1,
                 buffers_per_step=96),
    # BASELINE.json configs[0] shape (CPU-runnable case), many steps so the GPU has work
    "cfg1": dict(desc="int16 IQ @8 MS/s, N=1024, K=16 averaging, DC off (configs[0] shape, batched)",
                 kind=3, enob=12, dc=False, n=1024, K=16, fs=8_000_000, start=300e6, stop=0.0, win=5,
                 buffers_per_step=16 * 16384),
}


def workload_config(name: str, wl: dict, spectrum: bool, n_steps: int) -> dict:
    """The `config` object: the same keys and values on BOTH arms (the driver compares them), nothing that depends
    on the arm, the GPU count or a calibrated value -- those live under `run`."""
    return {"workload": name + ": " + wl["desc"], "fft_size": wl["n"], "averaging": wl["K"],
            "retune_steps": n_steps, "sample_kind": {1: "int8 IQ", 2: "int16 split", 3: "int16 IQ", 4: "fp32 IQ"}[wl["kind"]],
            "enob": wl["enob"], "dc_correction": bool(wl["dc"]), "sample_rate": wl["fs"],
            "window": {1: "hann", 5: "blackman-harris"}[wl["win"]], "spectrum_written": bool(spectrum),
            "threshold_rule": "99.9th percentile of candidate-bin dB on a fixed-seed sample, guard-banded",
            "l2_policy": "inputs larger than L2 (GPU arm: >= 0.8 GB of raw IQ per GPU per step)"}


def bytes_per_sample(kind: int) -> int:
    return {1: 2, 2: 4, 3: 4, 4: 8}[kind]


def algorithmic_bytes_per_sample(wl: dict, spectrum: bool) -> float:
    """B_in + S*4/K + (N/8 mask bytes + 4 count bytes) / (K*N)."""
    n, K = wl["n"], wl["K"]
    return bytes_per_sample(wl["kind"]) + (4.0 / K if spectrum else 0.0) + (n / 8.0 + 4.0) / (K * n)


# ---------------------------------------------------------------------------------------------
# synthetic IQ: noise sigma 0.05 FS/rail + <= 4 tones in the used band,
# amplitudes log-uniform in [-40, -3] dBFS, +0.01 FS DC, round-to-nearest, saturate.
# ---------------------------------------------------------------------------------------------
def synth_device(torch, wl: dict, n_buffers: int, seed: int, device) -> "torch.Tensor":
    n, kind, enob = wl["n"], wl["kind"], wl["enob"]
    gen = torch.Generator(device=device)
    gen.manual_seed(0x5CA77E2 + seed)
    fs_amp = 1.0 if kind == 4 else float((1 << (enob - 1)) - 1)
    use_w = int(0.75 * n / 2.0)
    t = torch.arange(n, device=device, dtype=torch.float32)
    chunks = []
    chunk = max(1, min(n_buffers, (1 << 22) // n))
    for c0 in range(0, n_buffers, chunk):
        c = min(chunk, n_buffers - c0)
        x = 0.05 * torch.randn((c, n, 2), generator=gen, device=device) + 0.01
        n_tones = torch.randint(0, 5, (c, 1), generator=gen, device=device)
        for k in range(4):
            i = torch.randint(n // 2 - use_w, n // 2 + use_w + 1, (c, 1), generator=gen, device=device)
            fbin = ((i + n // 2) % n).to(torch.float32)
            amp_db = -40.0 + 37.0 * torch.rand((c, 1), generator=gen, device=device)
            amp = torch.where(n_tones > k, 10.0 ** (amp_db / 20.0), torch.zeros_like(amp_db))
            ph = 2 * np.pi * torch.rand((c, 1), generator=gen, device=device)
            arg = 2 * np.pi * torch.remainder(fbin * t[None, :], float(n)) / n + ph
            x[..., 0] += amp * torch.cos(arg)
            x[..., 1] += amp * torch.sin(arg)
        if kind == 4:
            chunks.append(x.contiguous())
        else:
            lo, hi = (-128, 127) if kind == 1 else (-32768, 32767)
            q = torch.clamp(torch.round(x * fs_amp), max(lo, -fs_amp - 1), min(hi, fs_amp))
            q = q.to(torch.int8 if kind == 1 else torch.int16)
            if kind == 2:
                q = q.permute(0, 2, 1).contiguous()   # [c][2][n]: re block then im block
            chunks.append(q.contiguous())
    return torch.cat(chunks, 0)


def synth_host(wl: dict, n_buffers: int, seed: int) -> np.ndarray:
    from tests import synth
    return synth.make_buffers(wl["kind"], wl["n"], n_buffers, wl["enob"], seed)


# ---------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.max_mhz, self._halt = [], set(), None, threading.Event()
        self.power = []
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1e3)
                except Exception:
                    pass
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self) -> dict:
        self._halt.set()
        self.join(timeout=1.0)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples), "sm_mhz_min": min(self.samples) if self.samples else None,
                "power_w_median": float(np.median(self.power)) if self.power else None,
                "power_w_max": max(self.power) if self.power else None}


def nvml_index(torch, local_rank: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            pass
    return local_rank


# ---------------------------------------------------------------------------------------------
def cpu_baseline(wl: dict, window: np.ndarray, threshold: float, use_w: int, target_s: float, threads: int = 0):
    """Times the oracle's `faithful` restatement of the reference CPU path (process.cpp:272-310 with
    its per-buffer copies) on a bounded sample.  Returns dict for the JSON line."""
    import oracle as O
    K = wl["K"]
    n_buf = max(K, (2048 // K) * K)
    raw = synth_host(wl, n_buf, seed=991)
    threads = threads or O.hardware_threads()
    sec, _, _ = O.bench(raw, wl["n"], wl["fs"], wl["enob"], wl["kind"], wl["dc"], 1, threshold, window, use_w,
                        repeats=1, threads=threads, faithful=2)
    repeats = max(1, int(target_s / max(sec, 1e-6)))
    sec, hits, threads = O.bench(raw, wl["n"], wl["fs"], wl["enob"], wl["kind"], wl["dc"], 1, threshold, window,
                                 use_w, repeats=repeats, threads=threads, faithful=2)
    samples = n_buf * wl["n"] * repeats
    return {"value": samples / sec / 1e6, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{n_buf} synthetic buffers of the workload x {repeats} repeats "
                      f"({samples / 1e6:.0f} Msamples, {sec:.1f} s), oracle restatement of "
                      f"process.cpp:272-310 incl. the reference's per-buffer copies; FFTs run 8 per AVX vector per thread (bit-identical to the scalar plan, tuned-library speed)"}, sec


def cpu_fft_upper_bound(wl: dict, target_s: float = 3.0):
    """FFT-ONLY throughput of a tuned CPU library (scipy's pocketfft, complex64, all cores) on the workload's size:
    an upper bound for what the reference's FFTW plan could reach on this host -- the oracle's own radix FFT stands
    in for FFTW in `cpu_baseline`, and the reference's convert / window / dB / detect stages are NOT in this number
    (SURVEY.md section 8d)."""
    try:
        import scipy.fft as sfft
    except Exception as e:      # noqa: BLE001
        return {"unavailable": str(e)}
    n, nb = wl["n"], max(64, (32 << 20) // (wl["n"] * 8))
    rng = np.random.default_rng(5)
    x = (rng.standard_normal((nb, n)) + 1j * rng.standard_normal((nb, n))).astype(np.complex64)
    workers = os.cpu_count() or 1
    sfft.fft(x, axis=1, workers=workers)
    t0 = time.perf_counter()
    reps = 0
    while time.perf_counter() - t0 < target_s:
        y = sfft.fft(x, axis=1, workers=workers)
        reps += 1
    sec = time.perf_counter() - t0
    assert y.dtype == np.complex64
    return {"value": nb * n * reps / sec / 1e6, "unit": UNIT, "cores": workers, "library": "scipy.fft (pocketfft) complex64",
            "sample": f"{nb} x {n}-point forward FFTs x {reps} repeats ({sec:.1f} s), FFT only"}


def run_reference(args) -> None:
    """Reference arm: the reference's own CPU algorithm (oracle restatement; the reference binary cannot
    be built here -- FFTW3/VOLK/gr-fft/Boost absent) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle as O
    wl = WORKLOADS[args.workload]
    window = O.window_build(wl["win"], wl["n"])
    use_w = O.use_window(0.75, wl["n"])
    K = wl["K"]
    n_buf = max(K, (4096 // K) * K)
    raw = synth_host(wl, n_buf, seed=991)
    thr = calibrate_threshold(wl, window, use_w)
    threads = O.hardware_threads()
    # size a step at ~1 s of CPU work
    sec, _, _ = O.bench(raw, wl["n"], wl["fs"], wl["enob"], wl["kind"], wl["dc"], 1, thr, window, use_w,
                        repeats=1, threads=threads, faithful=2)
    repeats = max(1, int(1.0 / max(sec, 1e-6)))
    for _ in range(args.warmup):
        O.bench(raw, wl["n"], wl["fs"], wl["enob"], wl["kind"], wl["dc"], 1, thr, window, use_w,
                repeats=repeats, threads=threads, faithful=2)
    total = 0.0
    for _ in range(args.steps):
        s, _, _ = O.bench(raw, wl["n"], wl["fs"], wl["enob"], wl["kind"], wl["dc"], 1, thr, window, use_w,
                          repeats=repeats, threads=threads, faithful=2)
        total += s
    samples_per_step = n_buf * wl["n"] * repeats
    value = samples_per_step * args.steps / total / 1e6
    sample = (f"each step = {n_buf} synthetic buffers x {repeats} repeats ({samples_per_step / 1e6:.0f} Msamples) "
              f"through the oracle restatement of process.cpp:272-310 (reference copies kept; FFTs 8 per AVX vector, bit-identical to the scalar plan), {threads} threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload, wl, not args.no_spectrum,
                                  len(O.frequency_table(wl["fs"], wl["start"], wl["stop"]))),
        "run": {"threshold_db": thr, "threads": threads},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def calibrate_threshold_gpu(wl: dict, window: np.ndarray, use_w: int, device: int) -> float:
    """Product arm: threshold = the 99.9th percentile of candidate-bin dB over a fixed-seed sample, taken from the
    GPU path's own spectra (no oracle on this arm outside the cpu_baseline / parity-check legs), nudged away from
    any sample bin (guard band): a few hits per spectrum, as a scanner would run."""
    import scanner_b200 as S
    from tests import synth
    K = wl["K"]
    raw = synth_host(wl, 16 * K, seed=77)
    with S.SpectrumSense(wl["n"], wl["fs"], wl["enob"], 3.0e38, window, sample_kind=wl["kind"],
                         correct_dc_offset=wl["dc"], averaging=K, max_spectra=16, flags=S.OUT_SPECTRUM,
                         device=device) as ss:
        db = ss.process(raw, want_hits=False)["spectra_db"].astype(np.float64)
    return synth.guard_banded_threshold(db, wl["n"], use_w, quantile=0.999)


def calibrate_threshold(wl: dict, window: np.ndarray, use_w: int) -> float:
    """Reference arm: the same rule evaluated with the double-precision oracle."""
    import oracle as O
    from tests import synth
    K = wl["K"]
    raw = synth_host(wl, 16 * K, seed=77)
    truth = O.pipeline(raw, wl["n"], wl["fs"], wl["enob"], wl["kind"], wl["dc"], K, 0.0, window, use_w,
                       precision=1, want_f64=True)
    return synth.guard_banded_threshold(truth["spectra_db64"], wl["n"], use_w, quantile=0.999)


# ---------------------------------------------------------------------------------------------
def plugin_surface(wl: dict) -> dict:
    """The reference-facing C++ surface end to end: source thread(s) -> SampleQueue (raw bytes into the pinned slab) ->
    ProcessSamples workers -> scn_submit_gather/scn_collect, i.e. what replaces scan.cpp:211-238, driven by the
    `scan_b200 bench` tool on this workload's buffer shape.  Two operating points: the reference's own hand-off
    (ONE producer thread, one AppendSamples call and one memcpy per buffer) and the batched hand-off (several producer
    threads, 64 buffers per AppendSamplesBatch call = one HackRF USB transfer, hackRFSource.cpp:251-264)."""
    import re
    import subprocess
    tool = os.path.join(ROOT, "scanner_b200", "scan_b200")
    if not os.path.exists(tool):
        return {"unavailable": "scanner_b200/scan_b200 not built"}
    n, kind = wl["n"], wl["kind"]
    cores = os.cpu_count() or 1
    producers = max(1, min(4, cores // 4))       # 4 producers + 2 workers measured best on 16- and 24-core hosts

    def run(total, workers, max_batch, prod, append, linger):
        cmd = [tool, "bench", str(kind), str(n), str(wl["enob"]), "1" if wl["dc"] else "0", "4096", str(total),
               str(workers), str(max_batch), str(prod), str(append), str(linger)]
        try:
            out = subprocess.run(cmd, capture_output=True, text=True, timeout=180)
        except Exception as e:      # noqa: BLE001
            return {"unavailable": str(e)}
        m = re.search(r"plugin-surface throughput: ([0-9.]+) Msamples/s \((.*)\)", out.stdout)
        if not m:
            return {"unavailable": (out.stdout + out.stderr)[-300:]}
        return {"value": float(m.group(1)), "unit": UNIT, "detail": m.group(2)}
    single = run(max(20000, int(3.0e9 // n)), 2, 4096, 1, 1, 0)
    batched = run(max(20000, int(12.0e9 // n)), 2, 8192, producers, 64, 200)
    out = dict(batched)
    out["path"] = ("%d ReplaySource threads x AppendSamplesBatch(64 buffers) -> SampleQueue (pinned slab) -> "
                   "ProcessSamples (2 workers, scn_submit_gather from the slab) -> C ABI; K=1 per-buffer detection" % producers)
    out["single_producer_single_appends"] = single
    return out


def kernel_only(torch, S, kind: int, n: int, K: int, dc: bool, win: int, enob: int, device: int, peak: float,
                target_bytes: int = 512 << 20, reps: int = 5) -> dict:
    """Kernel-only throughput of one (kind, N, K) configuration on device-resident random IQ larger than L2:
    3 warm-up launches, `reps` timed ones (CUDA events on the launch stream, median).  per_config rows."""
    dev = torch.device("cuda", device)
    bps = bytes_per_sample(kind)
    nbuf = max(K, (target_bytes // (n * bps)) // K * K)
    ns = nbuf // K
    if kind == 4:
        raw = (0.05 * torch.randn((nbuf, n, 2), device=dev)).contiguous()
    else:
        amp = 20 if kind == 1 else 300
        raw = torch.randint(-amp, amp, (nbuf, n, 2), device=dev, dtype=torch.int8 if kind == 1 else torch.int16)
    w = S.window_build(win, n)
    ctx = S.SpectrumSense(n, 20_000_000, enob, 40.0, w, sample_kind=kind, correct_dc_offset=dc, averaging=K,
                          max_spectra=16, max_hits_per_spectrum=16, device=device)
    d_spec = torch.empty((ns, n), dtype=torch.float32, device=dev)
    d_mask = torch.empty((ns, n // 32), dtype=torch.int32, device=dev)
    d_cnt = torch.empty((ns,), dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream()

    def go():
        ctx.launch_device(raw.data_ptr(), ns, d_spec.data_ptr(), d_mask.data_ptr(), d_cnt.data_ptr(), 0, 0, st.cuda_stream)
    for _ in range(3):
        go()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); go(); e1.record(st); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    samples = nbuf * n
    bpsamp = bps + 4.0 / K + (n / 8 + 4) / (K * n)
    out = {"kind": {1: "int8 IQ", 3: "int16 IQ", 4: "fp32 IQ"}[kind], "fft_size": n, "averaging": K, "dc": dc,
           "kernel": ctx.kernel_name, "kernel_ms": ms, "msamples_per_s": samples / ms / 1e3,
           "bytes_per_sample": bpsamp, "hbm_gbs": samples * bpsamp / ms / 1e6,
           "hbm_frac": samples * bpsamp / ms / 1e6 / peak, "input_bytes": int(raw.numel() * raw.element_size())}
    ctx.close()
    del raw, d_spec, d_mask, d_cnt
    torch.cuda.empty_cache()
    return out


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--buffers-per-step", type=int, default=0, help="per-GPU dwell (buffers per retune step)")
    ap.add_argument("--no-spectrum", action="store_true", help="detections only (S=0 in SURVEY.md 8d)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip per_config / sustained / plugin_e2e (N=1 extras)")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N > 1: how the per-step records travel -- NVLink peer-memory windows (scn_exchange_*) or "
                         "an NCCL all-gather on a side stream")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--sustain-seconds", type=float, default=3.5)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    import scanner_b200 as S
    from scanner_b200 import binding as B

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    wl = dict(WORKLOADS[args.workload])
    if args.buffers_per_step:
        wl["buffers_per_step"] = args.buffers_per_step
    n, K, kind = wl["n"], wl["K"], wl["kind"]
    spectrum = not args.no_spectrum
    table = S.frequency_table(wl["fs"], wl["start"], wl["stop"])
    n_steps = len(table)
    window = S.window_build(wl["win"], n)
    use_w = S.use_window(0.75, n)
    thr = calibrate_threshold_gpu(wl, window, use_w, local_rank)

    # ---- shard: global units (spectra) in step-major order, contiguous range per rank ------------
    spectra_per_step_per_gpu = max(1, wl["buffers_per_step"] // K)
    units_per_step = spectra_per_step_per_gpu * world          # dwell grows with N (weak scaling)
    shard = S.plan_shard(n_steps, units_per_step, rank, world)
    first_unit, n_spectra = shard.first_unit, shard.n_units
    n_buffers = n_spectra * K
    samples_per_rank = n_buffers * n

    raw = synth_device(torch, wl, n_buffers, seed=1000 * rank + 1, device=dev)
    raw_bytes = raw.numel() * raw.element_size()
    assert raw_bytes == n_buffers * n * bytes_per_sample(kind)

    HIT_CAP = 16
    flags = (S.OUT_SPECTRUM if spectrum else 0) | S.OUT_HITS
    e2e_chunk = min(n_spectra, max(1, (64 << 20) // (n * bytes_per_sample(kind) * K)))
    ctx = S.SpectrumSense(n, wl["fs"], wl["enob"], thr, window, sample_kind=kind, correct_dc_offset=wl["dc"],
                          averaging=K, max_spectra=e2e_chunk, max_hits_per_spectrum=HIT_CAP, flags=flags,
                          device=local_rank, ticket_slots=3)
    words, rec_words = ctx.words, ctx.record_words
    d_spec = torch.empty((n_spectra, n), dtype=torch.float32, device=dev) if spectrum else None
    # masks / counts are TRIPLE-buffered.  The side-stream work of batch i (summarize + exchange) cannot get an SM while
    # the persistent fused kernel of batch i+1 owns them all, so it really runs at the NEXT boundary, next to the start
    # of batch i+2; with two buffers batch i+2 had to wait for it (it reuses batch i's buffers) and the whole side
    # stream was serialised into every step; with three it only waits for batch i-1's, which finished a step ago.
    NBUF = 3
    d_masks = [torch.empty((n_spectra, words), dtype=torch.int32, device=dev) for _ in range(NBUF)]
    d_counts = [torch.empty((n_spectra,), dtype=torch.int32, device=dev) for _ in range(NBUF)]
    d_mask, d_count = d_masks[0], d_counts[0]
    stream = torch.cuda.current_stream()
    sh = stream.cuda_stream
    d_merged = torch.zeros((n_steps, rec_words), dtype=torch.int32, device=dev)
    # A step = the fused kernel on the main stream; everything after it -- the per-retune-step records
    # (scn_summarize_steps) and, N > 1, their exchange -- runs on a high-priority side stream UNDER the fused kernel of
    # the next batch:
    #   "peer": publish(i) + merge(i-1), two tiny kernels over NVLink peer-memory windows, no rank ever waits for
    #           another (scn_exchange.cu);   "nccl": all-gather + merge.
    # The fused kernel hands out its tiles dynamically (WorkQueue), so the CTAs those small kernels displace at the
    # start of a batch cost their own run time, not a whole static share of the batch.
    # (Folding the records into the fused kernel with atomics was tried and removed: ~0.5 M atomics per batch on the
    # 2-3 L2 lines of the current step's record cost 7 % of the kernel, and the extra code another 7 % in registers.)
    xch = None
    if world > 1 and args.exchange == "peer":
        xch = S.open_record_exchange(local_rank, rank, world, n_steps, rec_words)
    d_recs = [torch.empty((n_steps, rec_words), dtype=torch.int32, device=dev) for _ in range(NBUF)]
    d_gathers = [torch.empty((world, n_steps, rec_words), dtype=torch.int32, device=dev) for _ in range(NBUF)] \
        if (world > 1 and xch is None) else None
    side = torch.cuda.Stream(device=dev, priority=-1)
    out_ready = [torch.cuda.Event() for _ in range(NBUF)]
    out_free = [torch.cuda.Event() for _ in range(NBUF)]
    step_no = [0]
    last_seq = [0]

    kernel_events = []

    def step(timed: bool) -> None:
        par = step_no[0] % NBUF
        step_no[0] += 1
        if step_no[0] > NBUF:
            stream.wait_event(out_free[par])                         # batch i-3's masks / counts have been summarised
        if timed:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
        ctx.launch_device(raw.data_ptr(), n_spectra, d_spec.data_ptr() if spectrum else 0, d_masks[par].data_ptr(),
                          d_counts[par].data_ptr(), 0, 0, sh)
        if timed:
            e1.record(stream)
            kernel_events.append((e0, e1))
        out_ready[par].record(stream)
        with torch.cuda.stream(side):
            side.wait_event(out_ready[par])
            ctx.summarize_steps(d_masks[par].data_ptr(), d_counts[par].data_ptr(), n_spectra, first_unit,
                                units_per_step, n_steps, d_recs[par].data_ptr(), side.cuda_stream)
            out_free[par].record(side)
            if xch is not None:
                # publish(i) + merge(i - 1) in one launch: the previous batch's rows landed long ago
                last_seq[0] = xch.step(d_recs[par].data_ptr(), d_merged.data_ptr(), side.cuda_stream)
            elif world > 1:
                S.gather_step_records(d_recs[par], world, out=d_gathers[par])   # NCCL all-gather, ~13 KB per rank
                ctx.merge_step_records(d_gathers[par].data_ptr(), world, n_steps, d_merged.data_ptr(),
                                       side.cuda_stream)

    def drain() -> None:
        """The last batch's exchange belongs to the timed region."""
        if xch is not None and last_seq[0]:
            with torch.cuda.stream(side):
                xch.merge(last_seq[0], d_merged.data_ptr(), side.cuda_stream)
        stream.wait_stream(side)

    def barrier() -> None:
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- parity spot check against the oracle (untimed) ------------------------------------------
    parity = None
    if rank == 0:
        import oracle as O
        # kernel only -- no collective here: the other ranks are not in this block
        ctx.launch_device(raw.data_ptr(), n_spectra, d_spec.data_ptr() if spectrum else 0, d_mask.data_ptr(),
                          d_count.data_ptr(), 0, 0, sh)
        torch.cuda.synchronize()
        ns = min(8, n_spectra)
        sample = raw[: ns * K].cpu().numpy()
        truth = O.pipeline(sample, n, wl["fs"], wl["enob"], kind, wl["dc"], K, thr, window, use_w, precision=1,
                           want_f64=True)
        got_mask = d_mask[:ns].cpu().numpy().view(np.uint32)
        # any bin within the guard band of the threshold may legitimately differ: exclude such spectra
        from tests import synth
        cand = truth["spectra_db64"][:, synth.candidate_bins(n, use_w)]
        safe = np.min(np.abs(cand - thr), axis=1) > 2e-3
        ok = np.array_equal(got_mask[safe], truth["hit_mask"][safe])
        db_ok = True
        if spectrum:
            got_db = d_spec[:ns].cpu().numpy().astype(np.float64)
            t64 = truth["spectra_db64"]
            strong = t64 >= (10 * np.log10(np.sqrt(np.mean((10 ** (t64 / 10)) ** 2, axis=1, keepdims=True))) - 10)
            db_ok = bool(np.abs(got_db - t64)[strong].max() < 1e-3)
        parity = {"spectra_checked": int(safe.sum()), "masks_equal": bool(ok), "db_within_1e-3": db_ok,
                  "hits_in_sample": int(truth["hit_count"].sum())}
        if not (ok and db_ok):
            raise SystemExit(f"bench.py: parity spot check failed: {parity}")

    # ---- device-resident timing (value) ----------------------------------------------------------
    # NVML is initialised BEFORE the barrier: anything rank 0 does between the barrier and its
    # first launch shows up as rank skew inside the other ranks' timed region.
    sampler = ClockSampler(nvml_index(torch, local_rank)) if rank == 0 else None
    for _ in range(args.warmup):
        step(False)
    if sampler:
        sampler.start()
    barrier()
    launches0 = ctx.launch_count
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(stream)
    h0 = time.perf_counter()
    for _ in range(args.steps):
        step(True)
    host_enqueue_ms = (time.perf_counter() - h0) * 1e3 / args.steps      # CPU time to enqueue one step (no sync inside)
    drain()
    t1.record(stream)
    barrier()
    clocks = sampler.stop() if sampler else None
    elapsed_ms = t0.elapsed_time(t1)
    launches = ctx.launch_count - launches0
    if xch is not None:
        launches += 2 * args.steps          # publish + merge per batch (scn_exchange kernels are not counted by the ctx)
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b in kernel_events]))
    starts = [a for a, _ in kernel_events] + [t1]
    step_ms = [round(starts[i].elapsed_time(starts[i + 1]), 4) for i in range(len(starts) - 1)]
    kern_ms_rank = kern_ms
    if world > 1:
        tmax = torch.tensor([elapsed_ms, kern_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        elapsed_ms, kern_ms = float(tmax[0]), float(tmax[1])
    total_samples = samples_per_rank * world
    ms_per_step = elapsed_ms / args.steps
    value = total_samples / (ms_per_step * 1e-3) / 1e6

    # ---- the exchanged records are RIGHT (untimed): merged == sum / OR over the ranks' partial records, and the
    # merged hit total == the all-reduced sum of the per-spectrum counts ---------------------------------------------
    records_check = None
    if world > 1:
        if xch is not None and xch.status() != 0:
            raise SystemExit(f"bench.py: record exchange timed out waiting for a peer (sequence {xch.status()})")
        local = d_recs[(step_no[0] - 1) % NBUF]
        sums = local[:, :2].clone().to(torch.int64)
        ors = local[:, 2:].clone()
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
        ors_all = torch.empty((world,) + tuple(ors.shape), dtype=ors.dtype, device=dev)    # NCCL has no bitwise OR
        dist.all_gather_into_tensor(ors_all.view(world * ors.shape[0], ors.shape[1]), ors.contiguous())
        for r in range(world):
            ors = ors | ors_all[r]
        hits_total = d_count.to(torch.int64).sum().reshape(1)
        dist.all_reduce(hits_total, op=dist.ReduceOp.SUM)
        ok_sum = bool(torch.equal(d_merged[:, :2].to(torch.int64), sums))
        ok_or = bool(torch.equal(d_merged[:, 2:], ors))
        ok_hits = int(d_merged[:, 0].to(torch.int64).sum()) == int(hits_total[0])
        ok_units = int(d_merged[:, 1].to(torch.int64).sum()) == units_per_step * n_steps
        records_check = {"merged_equals_sum_of_partials": ok_sum, "merged_masks_equal_or_of_partials": ok_or,
                         "merged_hit_total_equals_allreduced_counts": ok_hits, "hit_total": int(hits_total[0]),
                         "every_unit_accounted": ok_units, "exchange": args.exchange}
        if not (ok_sum and ok_or and ok_hits and ok_units):
            raise SystemExit(f"bench.py: merged step records are wrong on rank {rank}: {records_check}")

    # ---- roofline of the dominant (fused) kernel -------------------------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        with open(peaks_path) as f:
            peak, peak_src = float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    bps = algorithmic_bytes_per_sample(wl, spectrum)
    achieved = samples_per_rank * bps / (kern_ms * 1e-3) / 1e9
    info = ctx.kernel_info()
    # DRAM bytes per launch from the committed ncu --set full capture of this workload (profiles/), if any
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and spectrum:
        with open(tpath) as f:
            tj = json.load(f).get(args.workload)
        if tj and tj["samples_per_launch"] == samples_per_rank:
            traffic, traffic_src = tj["dram_bytes_read"] + tj["dram_bytes_write"], tj["source"]
    fp32 = None
    if os.path.exists(tpath):
        with open(tpath) as f:
            fj = (json.load(f).get(args.workload) or {}).get("fp32_pipe")
        if fj and traffic is not None:
            peak_lane = 148 * 128 * 1.965e9                       # FP32 lanes x SMs x max SM clock
            ach = samples_per_rank * fj["lane_cycles_per_sample"] / (kern_ms * 1e-3)
            fp32 = {"lane_cycles_per_sample": fj["lane_cycles_per_sample"], "achieved_lane_ops_per_s": ach,
                    "peak_lane_ops_per_s": peak_lane, "frac": ach / peak_lane, "source": fj["source"],
                    "note": "secondary roofline: 1-byte IQ is FP32-pipe bound (SURVEY.md H1); informational"}
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum)",
                "traffic_source": traffic_src, "algorithmic_bytes_per_launch": samples_per_rank * bps,
                "fp32_pipe": fp32, "kernel": ctx.kernel_name, "kernel_ms": kern_ms,
                "algorithmic_bytes_per_sample": bps, "peak_source": peak_src,
                "kernel_msamples_per_s": samples_per_rank / (kern_ms * 1e-3) / 1e6, "occupancy": info}

    # ---- end to end through scn_submit/scn_collect with pinned host buffers --------------------------
    e2e = None
    if not args.no_e2e:
        host_addr, host_view = B.alloc_pinned(raw_bytes)
        host_t = torch.from_numpy(host_view)
        host_t.copy_(raw.view(torch.uint8).reshape(-1).cpu())
        chunk_bytes = e2e_chunk * K * n * bytes_per_sample(kind)
        n_chunks = (n_spectra + e2e_chunk - 1) // e2e_chunk
        h_counts = np.empty(n_spectra, np.uint32)
        h_masks = np.empty((n_spectra, words), np.uint32)
        h_hits = np.zeros((n_spectra, HIT_CAP), S.hit_dtype)

        # bare-copy ceiling: the same chunks, pinned host -> device, three streams, nothing else; every rank at once
        cp_streams = [torch.cuda.Stream(device=dev) for _ in range(3)]
        d_land = [torch.empty(chunk_bytes, dtype=torch.uint8, device=dev) for _ in range(3)]

        def copy_pass() -> None:
            for c in range(n_chunks):
                nb = min(chunk_bytes, raw_bytes - c * chunk_bytes)
                with torch.cuda.stream(cp_streams[c % 3]):
                    d_land[c % 3][:nb].copy_(host_t[c * chunk_bytes: c * chunk_bytes + nb], non_blocking=True)
            torch.cuda.synchronize()
        copy_pass()
        barrier()
        c0 = time.perf_counter()
        for _ in range(3):
            copy_pass()
        copy_s = (time.perf_counter() - c0) / 3
        if world > 1:
            tt = torch.tensor([copy_s], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            copy_s = float(tt[0])
        h2d_ceiling_gbs = raw_bytes * world / copy_s / 1e9
        del d_land

        def e2e_step() -> int:
            inflight = []
            def collect_one():
                tk, first, cnt = inflight.pop(0)
                ctx.collect(tk, None, h_masks[first:first + cnt], h_counts[first:first + cnt],
                            h_hits[first:first + cnt], None)
            for c in range(n_chunks):
                first = c * e2e_chunk
                cnt = min(e2e_chunk, n_spectra - first)
                if len(inflight) == 3:
                    collect_one()
                inflight.append((ctx.submit(host_addr + c * chunk_bytes, cnt), first, cnt))
            while inflight:
                collect_one()
            return int(h_counts.sum())

        e2e_steps = max(1, min(args.steps, 10))
        for _ in range(2):
            e2e_step()
        barrier()
        w0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_hits = e2e_step()
        torch.cuda.synchronize()
        w1 = time.perf_counter()
        e2e_s = w1 - w0
        if world > 1:
            tt = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            e2e_s = float(tt[0])
        # no hit record was dropped: the cap holds for every spectrum of the batch (the C++ host re-runs a spectrum
        # that overflows, process.cpp; this leg has no such fallback, so it must not need one)
        max_hits = int(h_counts.max())
        if max_hits > HIT_CAP:
            raise SystemExit(f"bench.py: a spectrum has {max_hits} hits > record capacity {HIT_CAP}: e2e leg would drop hits")
        same_as_device = int(h_counts.sum()) == int(d_count.to(torch.int64).sum())
        d2h = n_spectra * (4 + 4 * words + HIT_CAP * 8)
        e2e_value = total_samples * e2e_steps / e2e_s / 1e6
        e2e = {"value": e2e_value, "unit": UNIT,
               "h2d_bytes_per_step": int(raw_bytes), "d2h_bytes_per_step": int(d2h), "steps": e2e_steps,
               "ms_per_step": e2e_s / e2e_steps * 1e3,
               "path": "scn_submit/scn_collect, pinned host raw buffers, 3 ticket slots x %d spectra; "
                       "D2H = hit counts + masks + <=%d hit records per spectrum; dB spectrum stays in HBM" % (e2e_chunk, HIT_CAP),
               "hits_per_step": e2e_hits, "max_hits_in_a_spectrum": max_hits, "hit_total_equals_device_path": same_as_device,
               "h2d_gbs": raw_bytes * world * e2e_steps / e2e_s / 1e9,
               "h2d_ceiling_gbs": h2d_ceiling_gbs,
               "h2d_ceiling_note": "bare pinned-host -> device copies of the same chunks on 3 streams, all ranks at "
                                   "once, max over ranks: the PCIe/host roofline of this leg",
               "frac_of_h2d_ceiling": (raw_bytes * world * e2e_steps / e2e_s / 1e9) / h2d_ceiling_gbs}
        B.free_pinned(host_addr)

    # ---- N = 1 extras: the other BASELINE configs kernel-only, a sustained-clock loop, the plugin surface ---------
    per_config = sustained = plugin = None
    if rank == 0 and world == 1 and not args.no_extras:
        del d_spec
        torch.cuda.empty_cache()
        rows = []
        for name in ("cfg1", "cfg3", "cfg4"):
            w_ = WORKLOADS[name]
            r = kernel_only(torch, S, w_["kind"], w_["n"], w_["K"], w_["dc"], w_["win"], w_["enob"], local_rank, peak)
            r["config"] = name
            rows.append(r)
        for n_ in (256, 1024, 4096, 8192, 16384, 65536):
            for kind_, dc_, enob_ in ((1, True, 8), (4, False, 0)):
                r = kernel_only(torch, S, kind_, n_, 1, dc_, 5, enob_, local_rank, peak)
                r["config"] = "cfg5 sweep"
                rows.append(r)
        per_config = rows
        # sustained: the headline kernel back to back for >= sustain-seconds with the SM clock / power sampled
        d_spec = torch.empty((n_spectra, n), dtype=torch.float32, device=dev) if spectrum else None
        s2 = ClockSampler(nvml_index(torch, local_rank))
        reps = max(10, int(args.sustain_seconds * 1e3 / kern_ms_rank))
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
        torch.cuda.synchronize()
        s2.start()
        evs[0].record(stream)
        for i in range(reps):
            ctx.launch_device(raw.data_ptr(), n_spectra, d_spec.data_ptr() if spectrum else 0, d_mask.data_ptr(),
                              d_count.data_ptr(), 0, 0, sh)
            evs[i + 1].record(stream)
        torch.cuda.synchronize()
        c2 = s2.stop()
        ts = np.array([evs[i].elapsed_time(evs[i + 1]) for i in range(reps)])
        q = max(1, reps // 4)
        sustained = {"seconds": float(ts.sum() / 1e3), "launches": reps,
                     "msamples_per_s": samples_per_rank / ts.mean() / 1e3,
                     "first_quarter_msamples_per_s": samples_per_rank / ts[:q].mean() / 1e3,
                     "last_quarter_msamples_per_s": samples_per_rank / ts[-q:].mean() / 1e3,
                     "clocks": c2}
        plugin = plugin_surface(wl)

    if rank == 0:
        cpu = None
        cpu2 = None
        cpu_fft = None
        if world == 1 and not args.no_cpu_baseline:
            cpu, _ = cpu_baseline(wl, window, thr, use_w, args.cpu_seconds)
            # the reference itself runs exactly two worker threads (scan.cpp:217)
            cpu2, _ = cpu_baseline(wl, window, thr, use_w, min(args.cpu_seconds, 4.0), threads=2)
            cpu_fft = cpu_fft_upper_bound(wl)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.workload, wl, spectrum, n_steps),
            "run": {"buffers_per_step_per_gpu": spectra_per_step_per_gpu * K,
                    "samples_per_gpu_per_step": samples_per_rank, "input_bytes_per_gpu": int(raw_bytes),
                    "l2_policy": "inputs_larger_than_l2" if raw_bytes > (126 << 20) else "inputs_fit_l2",
                    "sharding": ("contiguous (retune step, buffer) units per rank; per-step records exchanged over "
                                 + ("NVLink peer-memory windows (scn_exchange: publish(i) + merge(i-1), no rendezvous)"
                                    if xch is not None else "an NCCL all-gather on a side stream")) if world > 1 else "single GPU",
                    "threshold_db": thr},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "cpu_baseline": cpu, "cpu_baseline_2_threads": cpu2, "cpu_fft_only_upper_bound": cpu_fft,
            "parity": parity, "records_check": records_check, "step_ms": step_ms,
            "host_enqueue_ms_per_step": host_enqueue_ms,
            "per_config": per_config, "sustained": sustained, "plugin_e2e": plugin,
        }
        print(json.dumps(line), flush=True)
    if xch is not None:
        barrier()
        xch.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
