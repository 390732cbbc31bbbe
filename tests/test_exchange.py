"""scn_exchange_* (NVLink peer-memory exchange of per-step records) on ONE device: the protocol -- slots, sequence
numbers, flags, the merge rule -- does not care whether a peer's window is in another GPU's HBM, so two ranks that
both live on cuda:0 (connected in-process) exercise all of it on a single-GPU box.  The cross-GPU / CUDA-IPC
transport is covered by tests/test_multirank.py::test_two_ranks_nccl_gpu and asserted in every bench.py run."""
import numpy as np
import pytest

import scanner_b200 as S

pytestmark = pytest.mark.gpu


def merge_rule(parts):
    out = np.zeros_like(parts[0])
    out[:, :2] = np.sum([p[:, :2] for p in parts], axis=0)
    out[:, 2:] = np.bitwise_or.reduce([p[:, 2:] for p in parts], axis=0)
    return out


def test_single_rank_roundtrip():
    import torch
    n_steps, rw = 7, 10
    x = S.RecordExchange(0, 0, 1, n_steps, rw)
    rng = np.random.default_rng(3)
    d_out = torch.zeros((n_steps, rw), dtype=torch.int32, device="cuda")
    for batch in range(1, 7):
        rec = rng.integers(0, 2 ** 31, (n_steps, rw), dtype=np.int64).astype(np.uint32)
        d_rec = torch.from_numpy(rec.view(np.int32)).cuda()
        assert x.publish(d_rec.data_ptr()) == batch
        x.merge(batch, d_out.data_ptr())
        torch.cuda.synchronize()
        assert np.array_equal(d_out.cpu().numpy().view(np.uint32), rec)
    assert x.status() == 0
    x.close()


@pytest.mark.parametrize("world", [2, 5])
def test_ranks_on_one_device_publish_then_merge_previous(world):
    import torch
    n_steps, rw = 50, 66                                   # cfg2: 50 retune steps, N = 2048
    xs = [S.RecordExchange(0, r, world, n_steps, rw) for r in range(world)]
    S.RecordExchange.connect_local(xs)
    rng = np.random.default_rng(11)
    d_out = [torch.zeros((n_steps, rw), dtype=torch.int32, device="cuda") for _ in range(world)]
    history = {}
    for batch in range(1, 10):                             # > 2 x the 4 slots
        parts = [rng.integers(0, 2 ** 20, (n_steps, rw), dtype=np.int64).astype(np.uint32) for _ in range(world)]
        history[batch] = merge_rule(parts)
        for r in range(world):
            d_rec = torch.from_numpy(parts[r].view(np.int32)).cuda()
            assert xs[r].publish(d_rec.data_ptr()) == batch
            if batch > 1:
                xs[r].merge(batch - 1, d_out[r].data_ptr())   # every rank of batch - 1 published before this point
        torch.cuda.synchronize()
        if batch > 1:
            for r in range(world):
                assert np.array_equal(d_out[r].cpu().numpy().view(np.uint32), history[batch - 1]), (batch, r)
    for r in range(world):
        xs[r].merge(9, d_out[r].data_ptr())
    torch.cuda.synchronize()
    for r in range(world):
        assert np.array_equal(d_out[r].cpu().numpy().view(np.uint32), history[9])
        assert xs[r].status() == 0
    # contract: merge(s) must come before publish(s + 2); a stale or future sequence number is refused
    with pytest.raises(S.ScannerError):
        xs[0].merge(7, d_out[0].data_ptr())
    with pytest.raises(S.ScannerError):
        xs[0].merge(10, d_out[0].data_ptr())
    for x in xs:
        x.close()


def test_step_is_publish_plus_merge_of_the_previous_batch():
    """scn_exchange_step: one launch that publishes batch i and merges batch i - 1 (the steady-state call of bench.py)."""
    import torch
    world, n_steps, rw = 3, 50, 66
    xs = [S.RecordExchange(0, r, world, n_steps, rw) for r in range(world)]
    S.RecordExchange.connect_local(xs)
    rng = np.random.default_rng(17)
    d_out = [torch.full((n_steps, rw), -1, dtype=torch.int32, device="cuda") for _ in range(world)]
    want = {}
    for batch in range(1, 8):
        parts = [rng.integers(0, 2 ** 20, (n_steps, rw), dtype=np.int64).astype(np.uint32) for _ in range(world)]
        want[batch] = merge_rule(parts)
        for r in range(world):
            assert xs[r].step(torch.from_numpy(parts[r].view(np.int32)).cuda().data_ptr(), d_out[r].data_ptr()) == batch
        torch.cuda.synchronize()
        # rank r's step(batch) ran before ranks > r published `batch`, but every rank had published batch - 1
        for r in range(world):
            got = d_out[r].cpu().numpy().view(np.uint32)
            if batch == 1:
                assert np.all(got == 0xFFFFFFFF)          # nothing to merge yet: output untouched
            else:
                assert np.array_equal(got, want[batch - 1]), (batch, r)
    for r in range(world):
        xs[r].merge(7, d_out[r].data_ptr())
    torch.cuda.synchronize()
    for r in range(world):
        assert np.array_equal(d_out[r].cpu().numpy().view(np.uint32), want[7]) and xs[r].status() == 0
        xs[r].close()


def test_unconnected_exchange_refuses_to_publish():
    import torch
    x = S.RecordExchange(0, 0, 2, 4, 10)
    d = torch.zeros((4, 10), dtype=torch.int32, device="cuda")
    with pytest.raises(S.ScannerError):
        x.publish(d.data_ptr())
    x.close()
