"""Replica mode: independent utterance streams partitioned one engine per GPU.

Batch-1 decode of a 0.6 B model does not shard (five grid-wide dependencies per layer; over NVLink each would
cost microseconds), so N GPUs run N independent engines and there is NO collective on the data path.
``torch.distributed`` is used only to line the ranks up and to combine their clocks: whole-job throughput =
units processed by all ranks / max over ranks of the device time.
"""

from __future__ import annotations

from typing import Optional, Sequence

import torch


def assign(utterances: Sequence, world_size: int, rank: int) -> list:
    """Static round-robin partition of an utterance list: rank r serves utterances r, r + world, ..."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside [0, {world_size})")
    return [u for i, u in enumerate(utterances) if i % world_size == rank]


def combine(local_units: float, local_ms: float, group=None, device: Optional[torch.device] = None) -> tuple[float, float]:
    """(sum of units over ranks, max of elapsed ms over ranks); identity without an initialised process group."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(local_units), float(local_ms)
    dev = device if device is not None else torch.device("cpu")
    u = torch.tensor([float(local_units)], dtype=torch.float64, device=dev)
    t = torch.tensor([float(local_ms)], dtype=torch.float64, device=dev)
    dist.all_reduce(u, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(u[0]), float(t[0])


def throughput(local_units: float, local_ms: float, group=None, device: Optional[torch.device] = None) -> float:
    """Whole-job units per second."""
    units, ms = combine(local_units, local_ms, group, device)
    return units / (ms / 1000.0)
