"""Loader for tests/golden/reference_vectors.npz (outputs of the reference's own sources, see
tests/golden/make_golden.py)."""
import os
import re

import numpy as np

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_vectors.npz")


def load():
    return np.load(PATH, allow_pickle=False)


def scan_cases(g):
    for row in g["scan_cases"]:
        name, kind, n, fs, enob, dc, per_sweep, sweeps, win, mode, thr = row
        yield dict(name=str(name), kind=int(kind), n=int(n), fs=int(fs), enob=int(enob), dc=bool(int(dc)),
                   per_sweep=int(per_sweep), sweeps=int(sweeps), win=int(win), mode=int(mode), thr=float(thr),
                   raw=g[f"scan_{name}_raw"], freqs=g[f"scan_{name}_freqs"], text=str(g[f"scan_{name}_text"]))


def parse_hits(text):
    return [(int(m.group(1)), float(m.group(2)))
            for m in re.finditer(r"freq (\d+) power_db (-?[\d.]+|-?inf|-?nan)", text)]


def parse_time_domain(text):
    """[(sequence id, max dB, centre frequency, min dB)] from process.cpp:227-232 lines."""
    pat = r"Sequence\[(\d+)\]: Max signal (-?[\d.]+|-?inf) above threshold (-?[\d.]+) frequency (\d+), min (-?[\d.]+|-?inf)"
    return [(int(m.group(1)), float(m.group(2)), float(m.group(4)), float(m.group(5))) for m in re.finditer(pat, text)]


def accepted_range(case):
    """SampleQueue drops everything before the second scan-start marker (messageQueue.h:67-72)."""
    return case["per_sweep"], case["per_sweep"] * case["sweeps"]
