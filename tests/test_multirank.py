"""N > 1 plumbing.  CPU: world_size-2 gloo run of the shard plan + record all-gather against a
single-process result.  GPU (needs >= 2 devices): the same through the CUDA kernels, NCCL and the peer-memory exchange."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import scanner_b200 as S
from tests.conftest import ROOT

WORKER = r'''
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.environ["SCN_ROOT"])
import oracle as O
import scanner_b200 as S
from tests import synth

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
use_gpu = os.environ["SCN_BACKEND"] == "nccl"
if use_gpu:
    torch.cuda.set_device(rank)
dist.init_process_group(os.environ["SCN_BACKEND"])
n, kind, enob, n_steps, per_step = 1024, 1, 8, 5, 6          # 30 units -> 15 per rank: step 2 is split
window = S.window_build(5, n)
use_w = S.use_window(0.75, n)
raw_all = synth.make_buffers(kind, n, n_steps * per_step, enob, seed=4242)
thr = 14.0
words = n // 32
plan = S.plan_shard(n_steps, per_step, rank, world)
mine = raw_all[plan.first_unit: plan.first_unit + plan.n_units]
rec = np.zeros((n_steps, words + 2), np.uint32)
if use_gpu:
    dev = torch.device("cuda", rank)
    with S.SpectrumSense(n, 20_000_000, enob, thr, window, sample_kind=kind, correct_dc_offset=True,
                         max_spectra=plan.n_units, device=rank) as ss:
        d_raw = torch.from_numpy(mine.view(np.uint8).reshape(-1)).to(dev)
        d_mask = torch.zeros((plan.n_units, words), dtype=torch.int32, device=dev)
        d_cnt = torch.zeros((plan.n_units,), dtype=torch.int32, device=dev)
        d_rec = torch.zeros((n_steps, words + 2), dtype=torch.int32, device=dev)
        ss.launch_device(d_raw.data_ptr(), plan.n_units, 0, d_mask.data_ptr(), d_cnt.data_ptr())
        ss.summarize_steps(d_mask.data_ptr(), d_cnt.data_ptr(), plan.n_units, plan.first_unit, per_step,
                           n_steps, d_rec.data_ptr())
        parts = S.gather_step_records(d_rec, world)
        d_out = torch.zeros_like(d_rec)
        ss.merge_step_records(parts.data_ptr(), world, n_steps, d_out.data_ptr())
        torch.cuda.synchronize()
        merged = d_out.cpu().numpy().view(np.uint32)
        # the same records through the NVLink peer-memory exchange (scn_exchange.cu, CUDA IPC windows): three
        # batches so the slots rotate; publish(i) then merge(i - 1), as bench.py does
        xch = S.open_record_exchange(rank, rank, world, n_steps, words + 2)
        d_x = torch.zeros_like(d_rec)
        for batch in range(1, 4):
            assert xch.publish(d_rec.data_ptr()) == batch
            if batch > 1:
                xch.merge(batch - 1, d_x.data_ptr())
                torch.cuda.synchronize()
                assert np.array_equal(d_x.cpu().numpy().view(np.uint32), merged), ("exchange", rank, batch)
        xch.merge(3, d_x.data_ptr())
        torch.cuda.synchronize()
        assert np.array_equal(d_x.cpu().numpy().view(np.uint32), merged) and xch.status() == 0
        dist.barrier()
        xch.close()
else:
    res = O.pipeline(mine, n, 20_000_000, enob, kind, True, 1, thr, window, use_w, precision=0)
    for u in range(plan.n_units):
        s = plan.step_of_local(u)
        rec[s, 0] += res["hit_count"][u]
        rec[s, 1] += 1
        rec[s, 2:] |= res["hit_mask"][u]
    parts = S.gather_step_records(torch.from_numpy(rec.view(np.int32)), world).numpy().view(np.uint32)
    merged = np.zeros_like(rec)
    merged[:, :2] = parts[:, :, :2].sum(axis=0)
    merged[:, 2:] = np.bitwise_or.reduce(parts[:, :, 2:], axis=0)
# single-process truth
full = O.pipeline(raw_all, n, 20_000_000, enob, kind, True, 1, thr, window, use_w, precision=0)
want = np.zeros((n_steps, words + 2), np.uint32)
for u in range(n_steps * per_step):
    s = u // per_step
    want[s, 0] += full["hit_count"][u]; want[s, 1] += 1; want[s, 2:] |= full["hit_mask"][u]
assert np.array_equal(merged, want), (rank, merged[:, :2], want[:, :2])
assert want[:, 0].sum() > 0 and np.all(want[:, 1] == per_step)
assert list(plan.steps_touched()) == ([0, 1, 2] if rank == 0 else [2, 3, 4])
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def launch(backend: str, world: int = 2):
    port = free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), SCN_ROOT=ROOT, SCN_BACKEND=backend)
        procs.append(subprocess.Popen([sys.executable, "-c", WORKER], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT))
    outs = [p.communicate(timeout=300)[0].decode() for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {r}:\n{o}"
        assert f"rank {r} ok" in o


def test_shard_plan_properties():
    for n_steps, per_step in ((50, 4096), (133, 96), (24, 48), (1, 7), (5, 6)):
        for world in (1, 2, 4, 8):
            plans = [S.plan_shard(n_steps, per_step * world, r, world) for r in range(world)]
            assert plans[0].first_unit == 0
            assert sum(p.n_units for p in plans) == n_steps * per_step * world
            assert all(plans[r].first_unit + plans[r].n_units == plans[r + 1].first_unit for r in range(world - 1))
            assert max(p.n_units for p in plans) - min(p.n_units for p in plans) <= 1
            covered = sorted(set(s for p in plans for s in p.steps_touched()))
            assert covered == list(range(n_steps))


def test_two_ranks_gloo_cpu():
    launch("gloo")


@pytest.mark.gpu
def test_two_ranks_nccl_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    launch("nccl")
