"""Multi-GPU sharding of a portfolio: one process per GPU, options split into contiguous blocks,
no data-path collective; only the price gather goes through torch.distributed (NCCL over
NVLink on the GPU box, gloo in the CPU tests).  SURVEY.md 8(e): PDEs are independent
(reference src/Math/kwFd1d.cpp:61-136 touches one PDE's column only), so nothing else is exchanged.
"""
from __future__ import annotations

from typing import Callable, Tuple

import numpy as np


def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of rank; sizes differ by at most one."""
    return (n * rank) // world, (n * (rank + 1)) // world


def gather_prices(local, n_total: int, world: int, rank: int, dist, device=None):
    """All-gather variable-length price shards (torch tensors, f64) into the full vector, on every
    rank, in option order.  `local` lives on `device` (cuda for NCCL, cpu for gloo)."""
    import torch

    if world == 1:
        return local
    sizes = [shard_bounds(n_total, world, r)[1] - shard_bounds(n_total, world, r)[0] for r in range(world)]
    pad = max(sizes)
    buf = torch.full((pad,), float("nan"), dtype=torch.float64, device=local.device)
    buf[: local.shape[0]] = local
    out = torch.empty((world * pad,), dtype=torch.float64, device=local.device)
    dist.all_gather_into_tensor(out, buf)
    return torch.cat([out[r * pad: r * pad + sizes[r]] for r in range(world)])


def price_sharded(price_fn: Callable[[np.ndarray], Tuple[str, np.ndarray]], options: np.ndarray, world: int,
                  rank: int, dist=None, device="cpu") -> Tuple[str, np.ndarray]:
    """Every rank prices its block with `price_fn` (a Pricer.price) and receives all prices.
    Errors are made global: if any rank failed, every rank returns that rank's message."""
    import torch

    n = options.shape[0]
    lo, hi = shard_bounds(n, world, rank)
    err, p = price_fn(options[lo:hi]) if hi > lo else ("", np.empty(0))
    if p is None:
        p = np.full(hi - lo, np.nan)
    if world == 1:
        return err, p
    local = torch.from_numpy(np.ascontiguousarray(p)).to(device)
    full = gather_prices(local, n, world, rank, dist)
    flag = torch.tensor([1 if err else 0], dtype=torch.int32, device=device)
    flags = [torch.zeros_like(flag) for _ in range(world)]
    dist.all_gather(flags, flag)
    bad = [r for r in range(world) if int(flags[r].item())]
    if bad:
        msgs = [None] * world
        dist.all_gather_object(msgs, err)
        err = msgs[bad[0]]
    return err, full.cpu().numpy()
