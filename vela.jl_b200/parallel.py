"""Walker sharding across ranks (one process per GPU, torch.distributed).

The path shards by independent units: each walker's lnpost depends only on its own parameter row and
on the replicated read-only model (src/likelihood/posterior.jl:21-23 is a race-free parallel loop).
Ranks take contiguous walker blocks; the only collective is one all-gather of the per-walker results
(NCCL over NVLink on GPUs; gloo on CPU in the tests).
"""

from __future__ import annotations

from typing import Callable, Tuple

import numpy as np


def shard_range(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of `n` walkers owned by `rank` (block size ceil(n / world))."""
    per = (n + world - 1) // world
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def lnpost_sharded(evaluate: Callable[[np.ndarray], np.ndarray], paramss: np.ndarray, group=None,
                   device: str = "cpu") -> np.ndarray:
    """Every rank evaluates its block of `paramss` with `evaluate` and all ranks receive the full vector.

    `evaluate` is `Pulsar.lnpost_vectorized` (CUDA) in production; the CPU tests pass the oracle."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = int(np.shape(paramss)[0])
    lo, hi = shard_range(n, world, rank)
    mine = np.asarray(evaluate(np.asarray(paramss)[lo:hi]), dtype=np.float64) if hi > lo else np.zeros(0)
    if world == 1:
        return mine
    per = (n + world - 1) // world
    send = torch.full((per,), float("nan"), dtype=torch.float64, device=device)
    send[: hi - lo] = torch.from_numpy(mine).to(device)
    recv = torch.empty(per * world, dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(recv, send, group=group)
    return recv[:n].cpu().numpy()
