"""Multi-GPU plumbing: the stream-sharded (replica) mode — one process per GPU, no data-path collective — and the shard
arithmetic / handle exchange of the tensor-parallel single-stream mode.

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


def tp_shard(cfg: dict, rank: int, world: int) -> dict:
    """The shard of the temporal transformer that tensor-parallel rank `rank` of `world` owns — the same arithmetic as
    msx_model_load_gguf_tp (csrc/engine.cu): heads [h0, h1) for q/k/v rows, KV ring and out_proj columns; the hidden
    slice [f0, f1) of the gated MLP on 256-weight (super-block) boundaries when the width allows it, else on 32."""
    H, F = cfg["num_heads"], cfg["hidden"]
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank / world")
    if H % world:
        raise ValueError("num_heads must be divisible by the tensor-parallel world size")
    hl = H // world
    unit = 256 if F % 256 == 0 else 32
    nu = F // unit
    f0, f1 = (rank * nu // world) * unit, ((rank + 1) * nu // world) * unit
    if f1 <= f0:
        raise ValueError("hidden size too small for this world size")
    dh = cfg["dim"] // H
    return {"h0": rank * hl, "h1": (rank + 1) * hl, "adim": hl * dh, "f0": f0, "f1": f1,
            "in_proj_rows": [(s * cfg["dim"] + rank * hl * dh, s * cfg["dim"] + (rank + 1) * hl * dh) for s in range(3)],
            "linear_in_rows": [(f0, f1), (F + f0, F + f1)]}


def tp_exchange(dist, payload: bytes) -> list:
    """all-gather one small bytes object per rank in rank order (NCCL id broadcast uses broadcast_object_list; the
    64-byte IPC handles of msx_stream_tp_export use this)"""
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, payload)
    return out
