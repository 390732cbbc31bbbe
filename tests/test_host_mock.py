"""The HOST layer (SampleQueue, ProcessSamples, sources, scan_b200's wiring) against the reference's golden stdout
and recording files WITHOUT a GPU: the same C++ sources are linked against tests/mock_abi (the C ABI implemented with
the oracle -- test infrastructure, never shipped) instead of libscanner_b200.so.  What the GPU tests in
test_host_surface.py / test_hackrf_sweep.py / test_record.py check through the real library, these check for the
host logic alone -- including under ThreadSanitizer."""
import hashlib
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle as O                      # noqa: E402,F401  (builds oracle/_build/liboracle.so)
from tests import golden_util as GU     # noqa: E402

G = GU.load()
GH = np.load(os.path.join(ROOT, "tests", "golden", "hackrf_vectors.npz"), allow_pickle=False)
GR = np.load(os.path.join(ROOT, "tests", "golden", "record_vectors.npz"), allow_pickle=False)
ENV = dict(os.environ, TZ="UTC")
HOST = os.path.join(ROOT, "scanner_b200", "csrc", "host")
SRCS = [os.path.join(ROOT, "scanner_b200", "csrc", "tools", "scan_b200.cpp"),
        os.path.join(ROOT, "tests", "mock_abi", "mock_scanner_abi.cpp")] + \
       sorted(os.path.join(HOST, f) for f in os.listdir(HOST) if f.endswith(".cpp"))


def build(path, extra):
    lib = os.path.join(ROOT, "oracle", "_build")
    cmd = ["g++", "-O1", "-g", "-std=c++17", "-I" + os.path.join(ROOT, "include"), "-I" + HOST, *extra, "-o", path, *SRCS,
           "-L" + lib, "-loracle", "-lpthread", "-Wl,-rpath," + lib]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    return path


@pytest.fixture(scope="module")
def tool(tmp_path_factory):
    return build(str(tmp_path_factory.mktemp("mock") / "scan_mock"), [])


def replay(tool, case, tmp_path, threads=1, legacy=False, prefix=(), env=None):
    rp, fp = str(tmp_path / "raw.bin"), str(tmp_path / "freq.bin")
    np.ascontiguousarray(case["raw"]).tofile(rp)
    np.ascontiguousarray(case["freqs"], np.float64).tofile(fp)
    cmd = [*prefix, tool, "replay", str(case["kind"]), str(case["n"]), repr(float(case["fs"])), str(case["enob"]),
           "1" if case["dc"] else "0", repr(case["thr"]), str(case["win"]), str(case["mode"]), str(case["per_sweep"]), rp, fp,
           str(threads)] + (["legacy"] if legacy else [])
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env or ENV)
    assert r.returncode == 0, r.stderr[-2000:]
    return r


@pytest.mark.parametrize("case", [c for c in GU.scan_cases(G) if c["n"] >= 256], ids=lambda c: c["name"])
def test_host_layer_prints_what_the_reference_prints(tool, case, tmp_path):
    text, want = replay(tool, case, tmp_path).stdout, case["text"]
    if case["mode"] == 2:
        got_hits, want_hits = GU.parse_hits(text), GU.parse_hits(want)
        assert [f for f, _ in got_hits] == [f for f, _ in want_hits] and want_hits
        assert max(abs(a - b) for (_, a), (_, b) in zip(got_hits, want_hits)) < 1e-3 + 1e-6
    else:
        got_td, want_td = GU.parse_time_domain(text), GU.parse_time_domain(want)
        assert [(s, f) for s, _, f, _ in got_td] == [(s, f) for s, _, f, _ in want_td] and want_td
    strip = lambda t: [re.sub(r"power_db .*|Max signal .*|Start scan at .*", "", l) for l in t.splitlines()
                       if not re.match(r"Frequency \d+:|Starting source thread|Stopping source thread", l)]
    assert strip(text) == strip(want)


def test_two_workers_report_every_hit(tool, tmp_path):
    case = next(c for c in GU.scan_cases(G) if c["name"] == "i8_dc_2048")
    got = sorted(GU.parse_hits(replay(tool, case, tmp_path, threads=2).stdout))
    want = sorted(GU.parse_hits(case["text"]))
    assert [f for f, _ in got] == [f for f, _ in want]


def test_more_than_64_hits_falls_back_to_the_full_list(tool, tmp_path):
    from tests import synth
    case = dict(next(c for c in GU.scan_cases(G) if c["name"] == "i8_1024"))
    n = case["n"]
    window, use_w = O.window_build(case["win"], n), O.use_window(0.75, n)
    res = O.pipeline(case["raw"], n, case["fs"], case["enob"], case["kind"], case["dc"], 1, 0.0, window, use_w, precision=0)
    case["thr"] = float(np.float32(np.median(res["spectra_db"][:, synth.candidate_bins(n, use_w)])))
    res = O.pipeline(case["raw"], n, case["fs"], case["enob"], case["kind"], case["dc"], 1, case["thr"], window, use_w, precision=0)
    lo, hi = GU.accepted_range(case)
    assert res["hit_count"][lo:hi].min() > 64
    want = []
    for b in range(lo, hi):
        _, _, bins = O.detect(res["spectra_db"][b], use_w, 4, case["thr"])
        want += [O.hit_frequency(case["freqs"][b], case["fs"], n, int(i)) for i in bins]
    assert [f for f, _ in GU.parse_hits(replay(tool, case, tmp_path).stdout)] == want


def test_averaged_sweep_through_the_queue_terminates(tool):
    """K = 4 groups through SampleQueue::GetNextBatch (the wake-up regression) with two workers."""
    r = subprocess.run([tool, "synth", "1", "1024", "20000000", "8", "1", "30.0", "2400000000.0", "2450000000.0", "8", "3",
                        "7", "2", "4"], capture_output=True, text=True, timeout=120, env=ENV)
    assert r.returncode == 0, r.stderr[-2000:]
    m = re.search(r"buffers (\d+) hits (\d+) launches (\d+)", r.stderr)
    assert m and int(m.group(1)) == 2 * 3 * 8            # 2 accepted sweeps x 3 steps x 8 buffers


@pytest.mark.parametrize("threads", [1, 2])
def test_staging_copy_and_slab_submits_give_the_same_output(tool, tmp_path, threads):
    """Default: batches go to scn_submit_gather straight from the queue's slab (FIFO pool, address runs); with
    SCN_STAGING_COPY the consumer packs every batch into a staging buffer first.  The mock reads the submitted
    pointers only at collect time, so a message recycled before its ticket is collected would corrupt the output."""
    staged = dict(ENV, SCN_STAGING_COPY="1")
    for env in (ENV, staged):
        for name in ("i8_dc_2048", "i16split_256", "f32_hann_1024", "i16_dc_512"):
            case = next((c for c in GU.scan_cases(G) if c["name"] == name), None)
            if case is None:
                continue
            a = sorted(GU.parse_hits(replay(tool, case, tmp_path, threads=threads, env=env).stdout))
            b = sorted(GU.parse_hits(case["text"]))
            assert [f for f, _ in a] == [f for f, _ in b] and b
    # a long averaged sweep wraps the slab several times: most batches come straight from the slab, none is lost
    args = [tool, "synth", "1", "1024", "20000000", "8", "1", "30.0", "2400000000.0", "2450000000.0", "600", "4", "7",
            str(threads), "4"]
    r = subprocess.run(args, capture_output=True, text=True, timeout=300, env=ENV)
    assert r.returncode == 0, r.stderr[-2000:]
    m = re.search(r"buffers (\d+) hits (\d+) launches (\d+) zero-copy batches (\d+)", r.stderr)
    assert m and int(m.group(1)) == 3 * 3 * 600
    assert int(m.group(4)) > 0
    base = subprocess.run(args, capture_output=True, text=True, timeout=300, env=staged)
    m0 = re.search(r"buffers (\d+) hits (\d+) launches (\d+) zero-copy batches (\d+)", base.stderr)
    assert m0 and m0.group(1) == m.group(1) and m0.group(2) == m.group(2) and int(m0.group(4)) == 0


@pytest.mark.parametrize("append_batch", [3, 64])
def test_batched_appends_print_what_single_appends_print(tool, tmp_path, append_batch):
    """SampleQueue::AppendSamplesBatch == that many AppendSamples calls: same first-sweep drop, same scan-start lines,
    same sequence order, hence the reference's stdout (single worker: print order is sequence order)."""
    env = dict(ENV, SCN_APPEND_BATCH=str(append_batch))
    for case in GU.scan_cases(G):
        if case["name"].startswith("i16split"):
            continue                                         # the split layout has no batched form
        if case["n"] < 256:
            continue
        got = replay(tool, case, tmp_path, threads=1, env=env).stdout
        strip = lambda t: [re.sub(r"power_db .*|Max signal .*|Start scan at .*", "", l) for l in str(t).splitlines()
                           if not re.match(r"Frequency \d+:|Starting source thread|Stopping source thread", l)]
        assert strip(got) == strip(case["text"]), case["name"]


def test_many_producers_batched_appends_lose_nothing(tool):
    """Four producer threads x 64-buffer appends x two workers with a batch linger against one producer x
    single appends: every buffer is processed once (same buffer and hit totals; K = 1 detection is order free)."""
    def run(*extra):
        r = subprocess.run([tool, "bench", "1", "2048", "8", "1", "64", "6144", *extra], capture_output=True, text=True,
                           timeout=300, env=ENV)
        assert r.returncode == 0, r.stderr[-2000:]
        m = re.search(r"\((\d+) buffers .* (\d+) hits, (\d+) launches of which (\d+) straight", r.stdout)
        assert m, r.stdout
        return int(m.group(1)), int(m.group(2)), int(m.group(3)), int(m.group(4))
    base = run("1", "256", "1", "1", "0")
    many = run("2", "256", "4", "64", "200")
    assert base[0] == many[0] == 6144 and base[1] == many[1] and base[1] > 0
    assert many[3] > 0


def sweep_args(gpus, averaging=1, exchange="nccl", report=0, per_step=12, iterations=4):
    # int8 IQ, N = 1024, 2.40 - 2.52 GHz at 20 MS/s: 8 retune steps; threshold low enough for plenty of hits
    return ["1", "1024", "20000000", "8", "1", "17.0", "2400000000.0", "2520000000.0", str(per_step), str(iterations), "7",
            str(gpus), str(averaging), exchange, str(report)]


def scan_lines(text):
    return [re.sub(r"Start scan at .*", "Start scan at <wall clock>", l) for l in text.splitlines()
            if not re.match(r"(Starting|Stopped) process thread|Starting source thread|Stopping source thread|Frequency \d+:|Elapsed time", l)]


@pytest.mark.parametrize("averaging", [1, 4])
def test_sweep_processor_prints_the_single_gpu_output_on_any_gpu_count(tool, averaging):
    """SweepProcessor (one worker per GPU, retune steps split across them, ordered emitter) against ProcessSamples with
    one worker: the same lines in the same order for 1, 2, 3 and 8 GPUs (the mock ABI stands in for the devices), with
    either record exchange; and the merged per-step records account for every spectrum and hit of the last sweep."""
    base = subprocess.run([tool, "synth", *sweep_args(0, averaging)[:11], "1", str(averaging)], capture_output=True, text=True,
                          timeout=300, env=ENV)
    assert base.returncode == 0, base.stderr[-2000:]
    want = scan_lines(base.stdout)
    assert sum(l.startswith("freq ") for l in want) > 50 and sum(l.startswith("Start scan") for l in want) == 3
    total = int(re.search(r"buffers (\d+) hits (\d+)", base.stderr).group(1))
    for gpus, exchange in ((1, "nccl"), (2, "nccl"), (3, "peer"), (8, "nccl"), (8, "peer")):
        r = subprocess.run([tool, "sweep", *sweep_args(gpus, averaging, exchange)], capture_output=True, text=True,
                           timeout=300, env=ENV)
        assert r.returncode == 0, r.stderr[-2000:]
        assert scan_lines(r.stdout) == want, (gpus, exchange)
        m = re.search(r"buffers (\d+) hits (\d+) launches (\d+) sweeps (\d+) per-gpu((?: \d+)+) last-sweep-records hits (\d+) "
                      r"spectra (\d+)", r.stderr)
        assert m, r.stderr
        per_gpu = [int(x) for x in m.group(5).split()]
        assert int(m.group(1)) == total == sum(per_gpu) and len(per_gpu) == gpus
        assert int(m.group(4)) == 3                                      # the first sweep is dropped (messageQueue.h:67-72)
        if gpus in (2, 8):
            assert min(per_gpu) > 0 and max(per_gpu) - min(per_gpu) <= 3 * 12      # contiguous balanced step ranges
        assert int(m.group(7)) == 8 * 12 // averaging                    # every spectrum of the last sweep is in the records


def test_sweep_report_comes_from_the_merged_records(tool):
    """The per-sweep report lines are built from the EXCHANGED records: per step, hits == the number of `freq` lines the
    step printed in that sweep, whichever GPU processed it."""
    r = subprocess.run([tool, "sweep", *sweep_args(3, 1, "peer", report=1)], capture_output=True, text=True, timeout=300, env=ENV)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = scan_lines(r.stdout)
    sweeps, cur = [], None
    for l in lines:
        if l.startswith("Start scan"):
            cur = {"hits": {}, "report": {}}
            sweeps.append(cur)
        elif l.startswith("freq "):
            hz = int(l.split()[1])
            step = min(range(8), key=lambda s: abs(2407.5e6 + 15e6 * s - hz))
            cur["hits"][step] = cur["hits"].get(step, 0) + 1
        elif l.startswith("sweep "):
            f = l.split()
            cur["report"][int(f[3])] = (int(f[7]), int(f[9]))
    assert len(sweeps) == 3
    for sw in sweeps:
        assert sorted(sw["report"]) == list(range(8))
        for step in range(8):
            assert sw["report"][step] == (12, sw["hits"].get(step, 0)), (step, sw["report"][step], sw["hits"].get(step, 0))


def test_hackrf_sweep_replay(tool, tmp_path):
    n, fs, start, stop, thr, iterations, valid = [GH["sweep_params"][i] for i in range(7)]
    f = str(tmp_path / "stream.bin")
    GH["sweep_stream"].tofile(f)
    r = subprocess.run([tool, "hackrf", str(int(n)), str(int(fs)), repr(float(start)), repr(float(stop)), repr(float(thr)),
                        str(int(iterations)), str(int(valid)), f, "1"], capture_output=True, text=True, timeout=120, env=ENV)
    assert r.returncode == 0, r.stderr[-2000:]
    want = str(GH["sweep_text"])
    got_hits, want_hits = GU.parse_hits(r.stdout), GU.parse_hits(want)
    assert [h for h, _ in got_hits] == [h for h, _ in want_hits] and want_hits
    strip = lambda t: [re.sub(r"power_db .*", "", l) for l in t.splitlines()
                       if not l.startswith("interpolateSamples") and not re.match(r"(Starting|Stopped) process thread", l)]
    assert strip(r.stdout) == strip(want)


def run_record(tool, tmp_path, prefix=(), averaging=1, pre_post=None):
    n, fs, enob, kind, dc, per_sweep, pre, post = [int(x) for x in GR["params"][:8]]
    if pre_post is not None:
        pre, post = pre_post
    thr = float(GR["params"][8])
    rp, fp = str(tmp_path / "raw.bin"), str(tmp_path / "freq.bin")
    GR["raw"].tofile(rp)
    GR["freqs"].astype(np.float64).tofile(fp)
    r = subprocess.run([*prefix, tool, "record", str(kind), str(n), repr(float(fs)), str(enob), str(dc), repr(thr), "5",
                        str(per_sweep), rp, fp, str(tmp_path / "rec-"), str(pre), str(post), "1", str(averaging)],
                       capture_output=True, text=True, timeout=300, env=ENV)
    assert r.returncode == 0, r.stderr[-2000:]
    return r


def test_recording_is_bit_identical_to_the_reference(tool, tmp_path):
    r = run_record(tool, tmp_path)
    got = r.stdout.replace(str(tmp_path) + os.sep, "")
    pick = lambda t, pat: [l for l in str(t).splitlines() if re.match(pat, l)]
    assert pick(got, r"BeginWrite|EndWrite") == pick(GR["text"], r"BeginWrite|EndWrite")
    assert pick(got, r"Writing") == pick(GR["text"], r"Writing")
    files = sorted(f for f in os.listdir(tmp_path) if f.startswith("rec-"))
    want = [(str(a), int(b), str(c)) for a, b, c in GR["files"]]
    assert [f[len("rec-"):] for f in files] == [w[0] for w in want]
    for f, (_, size, sha) in zip(files, want):
        data = open(os.path.join(tmp_path, f), "rb").read()
        assert len(data) == size and hashlib.sha256(data).hexdigest() == sha


def test_recording_window_closes_with_averaging(tool, tmp_path):
    """K = 2, pre = post = 0 (the reference's defaults for multi-frequency scans): only ids 0, 2, 4, ... reach
    ProcessWrite, so `sequenceId == end` (end = trigger + 1, odd) never holds; the window must still close, at the
    end id, when the first later group arrives -- not run on to the end of the stream."""
    r = run_record(tool, tmp_path, averaging=2, pre_post=(0, 0))
    got = r.stdout.replace(str(tmp_path) + os.sep, "")
    marks = [l for l in got.splitlines() if re.match(r"BeginWrite|EndWrite", l)]
    begins = [int(l.rsplit(" ", 1)[1]) for l in marks if l.startswith("BeginWrite")]
    ends = [int(l.split()[1]) for l in marks if l.startswith("EndWrite")]
    # loud buffers carry sequence ids 5, 6, 7, 16, 25 -> groups (4,5) (6,7) | (16,17) | (24,25) trigger
    assert begins == [4, 16, 24], marks
    assert ends == [7, 17, 25], marks
    assert [m.split()[0].rstrip(":") for m in marks] == ["BeginWrite", "EndWrite"] * 3, marks
    files = sorted(f for f in os.listdir(tmp_path) if f.startswith("rec-"))
    # windows [4, 7), [16, 17), [24, 25) in sequence ids: 3 + 1 + 1 messages of 2048 fftwf_complex
    assert [os.path.getsize(os.path.join(tmp_path, f)) for f in files] == [3 * 2048 * 8, 2048 * 8, 2048 * 8]


def test_host_layer_is_clean_under_thread_sanitizer(tmp_path):
    """Two workers + the producer(s) + (record mode) the writer thread, with ThreadSanitizer watching the host code."""
    probe = subprocess.run(["g++", "-fsanitize=thread", "-x", "c++", "-", "-o", str(tmp_path / "probe")],
                           input="int main(){return 0;}", capture_output=True, text=True)
    if probe.returncode != 0:
        pytest.skip("no ThreadSanitizer runtime")
    tsan = build(str(tmp_path / "scan_tsan"), ["-fsanitize=thread"])
    prefix = ("setarch", "x86_64", "-R")                  # TSan wants a fixed address-space layout
    env_ok = subprocess.run([*prefix, "true"], capture_output=True).returncode == 0
    prefix = prefix if env_ok else ()
    case = next(c for c in GU.scan_cases(G) if c["name"] == "i8_dc_2048")
    r1 = replay(tsan, case, tmp_path, threads=2, prefix=prefix)
    r2 = run_record(tsan, tmp_path, prefix=prefix)
    # four producer threads appending 64-buffer batches + two workers submitting straight from the slab + linger
    r3 = subprocess.run([*prefix, tsan, "bench", "1", "2048", "8", "1", "64", "3000", "2", "256", "4", "64", "200"],
                        capture_output=True, text=True, timeout=600, env=ENV)
    assert r3.returncode == 0, r3.stderr[-2000:]
    # SweepProcessor: router + 3 device workers + ordered emitter + sweep barrier
    r4 = subprocess.run([*prefix, tsan, "sweep", *sweep_args(3, 2, "peer", report=1)], capture_output=True, text=True,
                        timeout=600, env=ENV)
    assert r4.returncode == 0, r4.stderr[-2000:]
    for r in (r1, r2, r3, r4):
        assert "ThreadSanitizer" not in r.stderr and "ThreadSanitizer" not in r.stdout, (r.stderr + r.stdout)[-4000:]
