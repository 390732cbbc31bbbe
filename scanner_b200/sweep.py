"""Multi-GPU plumbing for a sharded sweep: which (retune step, buffer) units a rank owns and the
exchange of per-retune-step detection records.

The path shards with no data-path collective: every retune step of the
FrequencyTable -- and every buffer within it -- is independent (process.cpp:279-309 carries nothing
between messages), so rank r simply owns a contiguous range of the sweep's units in step-major
order.  The only exchange is one all-gather of the small per-step records
(scn_summarize_steps output: [hit total, spectra contributing, OR-mask words]) per batch; NCCL
over NVLink on GPUs, any torch.distributed backend (gloo in the CPU tests) otherwise.
No arithmetic of the hot path lives here.
"""
from __future__ import annotations

from dataclasses import dataclass

from . import binding as B


@dataclass(frozen=True)
class ShardPlan:
    n_steps: int            # retune steps in the frequency table
    units_per_step: int     # spectra per retune step, summed over all ranks
    rank: int
    world: int
    first_unit: int         # this rank's first global unit (step-major)
    n_units: int            # units this rank owns

    @property
    def total_units(self) -> int:
        return self.n_steps * self.units_per_step

    def steps_touched(self) -> range:
        """Retune steps this rank holds at least one unit of (contiguous)."""
        if self.n_units == 0:
            return range(0)
        return range(self.first_unit // self.units_per_step,
                     (self.first_unit + self.n_units - 1) // self.units_per_step + 1)

    def step_of_local(self, local_index: int) -> int:
        return (self.first_unit + local_index) // self.units_per_step


def plan_shard(n_steps: int, units_per_step: int, rank: int, world: int) -> ShardPlan:
    """Contiguous, balanced (sizes differ by at most one unit) split of n_steps * units_per_step units;
    the split arithmetic is scn_shard_steps so C++ hosts and Python agree."""
    first, end = B.shard_steps(n_steps * units_per_step, rank, world)
    return ShardPlan(n_steps, units_per_step, rank, world, first, end - first)


def gather_step_records(records, world: int, out=None):
    """All-gathers each rank's [n_steps, record_words] partial records into [world, n_steps, record_words]
    on the tensor's own device/backend.  world == 1: returns records[None] without any collective."""
    if world == 1:
        return records.unsqueeze(0)
    import torch
    import torch.distributed as dist
    if out is None:
        out = torch.empty((world,) + tuple(records.shape), dtype=records.dtype, device=records.device)
    # concatenation-shaped view: accepted by both the NCCL and the gloo backend
    dist.all_gather_into_tensor(out.view((world * records.shape[0],) + tuple(records.shape[1:])), records.contiguous())
    return out


def open_record_exchange(device: int, rank: int, world: int, n_steps: int, record_words: int):
    """NVLink peer-memory exchange of the per-step records (scn_exchange.cu) for one-process-per-GPU jobs: every
    rank creates its window, the 64-byte CUDA IPC handles travel over torch.distributed (plumbing), every rank
    maps its peers' windows.  world == 1 needs no handles."""
    x = B.RecordExchange(device, rank, world, n_steps, record_words)
    if world > 1:
        import torch.distributed as dist
        handles = [None] * world
        dist.all_gather_object(handles, x.handle())
        x.connect_ipc(handles)
        dist.barrier()          # every window is mapped before anybody publishes into it
    return x
