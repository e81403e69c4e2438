"""Multi-GPU plumbing for the stream-sharded (replica) mode: one process per GPU, no data-path collective.

The LM decode path shards by independent conversation streams: every rank holds a full
replica of the weights and serves its own streams; torch.distributed is only used for the start/stop
barriers of a measurement and for the max-over-ranks reduction of device times.
"""
from __future__ import annotations


def assign_streams(n_streams: int, world: int) -> list[list[int]]:
    """round-robin stream ids to ranks (config 5 of BASELINE.json: 64 streams -> 8 per GPU on 8 GPUs)"""
    if world <= 0:
        raise ValueError("world must be positive")
    return [[s for s in range(n_streams) if s % world == r] for r in range(world)]


def aggregate_throughput(units_per_rank, elapsed_ms_per_rank) -> float:
    """whole-job units/s = total units / max-over-ranks time"""
    t = max(elapsed_ms_per_rank)
    return sum(units_per_rank) / (t * 1e-3)


def reduce_max_ms(local_ms: float, dist=None, device=None) -> float:
    """max over ranks of a device-measured time; dist = torch.distributed (initialised) or None"""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(local_ms)
    import torch
    t = torch.tensor([local_ms], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])
