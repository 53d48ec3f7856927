"""Batch sharding across GPUs: ciphertexts are independent, keys are replicated, so
each rank owns a contiguous slice of the batch and no collective touches the data path.
torch.distributed is used only for the barrier and the max-over-ranks of the device time."""
from __future__ import annotations


def shard_range(batch: int, rank: int, world: int) -> tuple[int, int]:
    """[start, stop) of the ciphertexts rank `rank` owns; sizes differ by at most one."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(batch, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def max_over_ranks(value: float, device=None) -> float:
    """MAX all-reduce of a scalar (device time in ms) over the default process group."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def aggregate_rate(units_per_rank: int, world: int, ms_max: float) -> float:
    """Whole-job throughput: units all ranks processed / slowest rank's time."""
    return world * units_per_rank / (ms_max * 1e-3)
