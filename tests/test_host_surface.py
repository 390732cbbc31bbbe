"""The C++ plugin surface (SignalSource / SampleQueue / ProcessSamples / ProcessInterface /
SampleBuffer under scanner_b200/csrc/host).  CPU: plumbing selftest.  GPU: scan_b200 (source ->
queue -> GPU ProcessSamples) must print what the reference prints for the same raw buffers."""
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

from tests import golden_util as GU
from tests.conftest import ROOT

G = GU.load()
TOOL = os.path.join(ROOT, "scanner_b200", "scan_b200")
SELFTEST = os.path.join(ROOT, "scanner_b200", "host_selftest")


def test_host_selftest_cpu():
    out = subprocess.run([SELFTEST], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=60)
    assert out.returncode == 0, out.stdout.decode()
    assert b"host_selftest ok" in out.stdout


def run_replay(case, threads=1, legacy=False):
    with tempfile.TemporaryDirectory() as d:
        rp, fp = os.path.join(d, "raw.bin"), os.path.join(d, "freq.bin")
        np.ascontiguousarray(case["raw"]).tofile(rp)
        np.ascontiguousarray(case["freqs"], np.float64).tofile(fp)
        cmd = [TOOL, "replay", str(case["kind"]), str(case["n"]), repr(float(case["fs"])), str(case["enob"]),
               "1" if case["dc"] else "0", repr(case["thr"]), str(case["win"]), str(case["mode"]),
               str(0 if legacy else case["per_sweep"]), rp, fp, str(threads)] + (["legacy"] if legacy else [])
        out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120)
        assert out.returncode == 0, out.stderr.decode()
        return out.stdout.decode()


@pytest.mark.gpu
@pytest.mark.parametrize("case", [c for c in GU.scan_cases(G) if c["n"] >= 256], ids=lambda c: c["name"])
def test_scan_b200_prints_what_the_reference_prints(case):
    text = run_replay(case)
    want = case["text"]
    if case["mode"] == 2:
        got_hits, want_hits = GU.parse_hits(text), GU.parse_hits(want)
        assert [f for f, _ in got_hits] == [f for f, _ in want_hits] and want_hits
        assert max(abs(a - b) for (_, a), (_, b) in zip(got_hits, want_hits)) < 1e-3 + 1e-6
    else:
        got_td, want_td = GU.parse_time_domain(text), GU.parse_time_domain(want)
        assert [(s, f) for s, _, f, _ in got_td] == [(s, f) for s, _, f, _ in want_td] and want_td
        assert all(abs(a[1] - b[1]) < 1e-5 and abs(a[3] - b[3]) < 1e-5 for a, b in zip(got_td, want_td))
    # same line structure: thread banners and one "Start scan at" per accepted sweep
    # (the golden run drives the reference's queue + ProcessSamples without a SignalSource, so the
    #  source-side banners -- table dump, "Starting source thread" -- are not part of it)
    strip = lambda t: [re.sub(r"power_db .*|Max signal .*|Start scan at .*", "", l) for l in t.splitlines()
                       if not re.match(r"Frequency \d+:|Starting source thread|Stopping source thread", l)]
    assert strip(text) == strip(want)


@pytest.mark.gpu
def test_legacy_samplebuffer_visitor_path():
    case = next(c for c in GU.scan_cases(G) if c["name"] == "i16_dc_512")
    text = run_replay(case, legacy=True)            # SampleBuffer has no first-sweep drop
    freqs = [int(m.group(1)) for m in re.finditer(r"freq (\d+)", text)]
    want = [f for f, _ in GU.parse_hits(case["text"])]
    assert len(freqs) > len(want) and freqs[-len(want):] == want


@pytest.mark.gpu
def test_synthetic_source_sweep_two_workers():
    # 3 sweeps x 4 steps x 8 buffers through 2 consumer threads (the reference's thread count)
    out = subprocess.run([TOOL, "synth", "1", "2048", "20000000", "8", "1", "30.0", "2.4e9", "2.46e9", "8", "3", "7", "2"],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120)
    assert out.returncode == 0, out.stderr.decode()
    m = re.search(r"buffers (\d+) hits (\d+) launches (\d+)", out.stderr.decode())
    assert m and int(m.group(1)) == 2 * 4 * 8 and int(m.group(3)) >= 1
    text = out.stdout.decode()
    assert text.count("Start scan at") == 2 and "Elapsed time" in text
    assert len(GU.parse_hits(text)) == int(m.group(2))
    # K = 4 averaging: 64 buffers -> 16 spectra
    out = subprocess.run([TOOL, "synth", "1", "2048", "20000000", "8", "1", "30.0", "2.4e9", "2.46e9", "8", "3", "7", "1", "4"],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120)
    assert out.returncode == 0, out.stderr.decode()
    m = re.search(r"buffers (\d+) hits (\d+) launches (\d+)", out.stderr.decode())
    assert m and int(m.group(1)) == 64


@pytest.mark.gpu
def test_hit_record_overflow_falls_back_to_full_list():
    """More than 64 hits in a buffer: ProcessSamples re-runs that spectrum at full capacity, so the printed
    list is still complete and ordered (checked against the oracle's detections)."""
    import oracle as O
    from tests import synth
    case = dict(next(c for c in GU.scan_cases(G) if c["name"] == "i8_1024"))
    n, kind = case["n"], case["kind"]
    window = O.window_build(case["win"], n)
    use_w = O.use_window(0.75, n)
    truth = O.pipeline(case["raw"], n, case["fs"], case["enob"], kind, case["dc"], 1, 0.0, window, use_w,
                       precision=1, want_f64=True)
    case["thr"] = synth.guard_banded_threshold(truth["spectra_db64"], n, use_w, guard=2e-3, quantile=0.6)
    res = O.pipeline(case["raw"], n, case["fs"], case["enob"], kind, case["dc"], 1, case["thr"], window, use_w,
                     precision=1)
    lo, hi = GU.accepted_range(case)
    assert res["hit_count"][lo:hi].min() > 64
    want = []
    for b in range(lo, hi):
        _, _, bins = O.detect(res["spectra_db"][b], use_w, 4, case["thr"])
        want += [O.hit_frequency(case["freqs"][b], case["fs"], n, int(i)) for i in bins]
    got = [f for f, _ in GU.parse_hits(run_replay(case))]
    assert got == want


# ---- SweepProcessor: one worker per GPU inside one process (csrc/host/sweepProcessor.cpp) -----------------------------
SWEEP = ["1", "1024", "20000000", "8", "1", "17.0", "2400000000.0", "2520000000.0", "12", "4", "7"]   # 8 retune steps


def _scan_lines(text):
    return [re.sub(r"Start scan at .*", "Start scan at <wall clock>", l) for l in text.splitlines()
            if not re.match(r"(Starting|Stopped) process thread|Starting source thread|Stopping source thread|"
                            r"Frequency \d+:|Elapsed time|NCCL version", l)]     # (NCCL announces itself on stdout)


def _run(args):
    out = subprocess.run([TOOL, *args], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)
    assert out.returncode == 0, out.stderr.decode()[-2000:]
    return out.stdout.decode(), out.stderr.decode()


@pytest.mark.gpu
@pytest.mark.parametrize("exchange", ["nccl", "peer"])
def test_sweep_processor_equals_process_samples_on_every_gpu_count(exchange):
    """`scan_b200 sweep` (SweepProcessor: steps split across GPUs, in-process ncclCommInitAll + ncclAllGather or the NVLink
    peer-memory windows for the per-sweep records) prints exactly what `scan_b200 synth` (ProcessSamples, one worker)
    prints, for 1 GPU and for every GPU count the box offers; the merged records account for the whole last sweep."""
    import torch
    want, err0 = _run(["synth", *SWEEP, "1", "1"])
    want = _scan_lines(want)
    assert sum(l.startswith("freq ") for l in want) > 100
    hits = int(re.search(r"buffers (\d+) hits (\d+)", err0).group(2))
    for gpus in sorted({1, min(2, torch.cuda.device_count()), torch.cuda.device_count()}):
        got, err = _run(["sweep", *SWEEP, str(gpus), "1", exchange, "0"])
        assert _scan_lines(got) == want, gpus
        m = re.search(r"buffers (\d+) hits (\d+) launches (\d+) sweeps (\d+) per-gpu((?: \d+)+) last-sweep-records hits (\d+) "
                      r"spectra (\d+)", err)
        assert m and int(m.group(2)) == hits and int(m.group(4)) == 3 and int(m.group(7)) == 8 * 12
        per_gpu = [int(x) for x in m.group(5).split()]
        assert len(per_gpu) == gpus and min(per_gpu) > 0


@pytest.mark.gpu
def test_batched_appends_from_several_producers_lose_nothing_on_the_gpu():
    """The ingest path at speed: 4 producer threads x AppendSamplesBatch(64) -> ring slab -> 2 workers submitting the slab
    runs with scn_submit_gather and reading results in place, against 1 producer x single AppendSamples: every buffer is
    processed exactly once (same buffer and hit totals; K = 1 detection does not depend on order or batch shape)."""
    def run(*extra):
        out, _ = _run(["bench", "1", "2048", "8", "1", "64", "61440", *extra])
        m = re.search(r"\((\d+) buffers .* (\d+) hits, (\d+) launches of which (\d+) straight", out)
        assert m, out
        return int(m.group(1)), int(m.group(2)), int(m.group(3)), int(m.group(4))
    base = run("1", "1024", "1", "1", "0")
    many = run("2", "4096", "4", "64", "200")
    assert base[0] == many[0] == 61440 and base[1] == many[1] and base[1] > 0
    assert many[3] == many[2] > 0                       # every launch came straight from the pinned slab
    staged = subprocess.run([TOOL, "bench", "1", "2048", "8", "1", "64", "61440", "2", "4096", "4", "64", "200"],
                            stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300, env=dict(os.environ, SCN_STAGING_COPY="1"))
    assert staged.returncode == 0
    m = re.search(r"\((\d+) buffers .* (\d+) hits, (\d+) launches of which (\d+) straight", staged.stdout.decode())
    assert m and int(m.group(2)) == base[1] and int(m.group(4)) == 0
