"""Multi-GPU sharding of the batch path: fields of view are independent, so they are dealt
round-robin to the ranks (FOV i -> rank i mod G) and nothing crosses ranks on
the data path.  ``torch.distributed`` is used only to gather the per-cell tables (host tensors)
and for the max-over-ranks timing of the benchmark; the backend may be NCCL (GPU box) or gloo
(CPU tests)."""

from __future__ import annotations

from typing import Any

import numpy as np


def shard_indices(n_fov: int, rank: int, world_size: int) -> np.ndarray:
    """Global FOV indices processed by ``rank``: i with i mod world_size == rank."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    return np.arange(rank, n_fov, world_size, dtype=np.int64)


def gather_fov_results(local: dict[int, Any], n_fov: int, dist=None, dst: int = 0) -> list[Any] | None:
    """Collect ``{global FOV index: result}`` from every rank onto ``dst`` in FOV order.

    ``dist`` is the initialised ``torch.distributed`` module (None = single process).  Results
    are small per-cell tables living on the host, so an object gather is all that is needed."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        parts = [local]
    else:
        parts = [None] * dist.get_world_size() if dist.get_rank() == dst else None
        dist.gather_object(local, parts, dst=dst)
        if parts is None:
            return None
    merged: dict[int, Any] = {}
    for part in parts:
        for idx, value in part.items():
            if idx in merged:
                raise ValueError(f"FOV {idx} was processed by two ranks")
            merged[idx] = value
    missing = [i for i in range(n_fov) if i not in merged]
    if missing:
        raise ValueError(f"FOVs {missing[:5]}... were not processed by any rank")
    return [merged[i] for i in range(n_fov)]


def max_over_ranks(seconds: float, dist=None, device=None) -> float:
    """Slowest rank's time (what a multi-GPU throughput number must be divided by)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(seconds)
    import torch

    t = torch.tensor([seconds], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])
