"""Multi-GPU data parallelism for the solve path.

Scenarios are independent NLPs, so the batch is split into contiguous shards, one process and
one handle per GPU, with NO collective during the solve; the only communication is the final
gather of (u0, cost, status, iters) = 32 bytes per scenario.
"""
from __future__ import annotations

import numpy as np


def shard_range(B: int, world: int, rank: int) -> tuple[int, int]:
    """Rank r owns [r*ceil(B/G), min(B,(r+1)*ceil(B/G)))."""
    per = -(-B // world)
    lo = min(B, rank * per)
    return lo, min(B, lo + per)


def balanced_permutation(B: int, seed: int = 0) -> np.ndarray:
    """Fixed-seed shuffle applied before sharding: iteration counts vary per scenario, shuffling
    keeps the per-GPU work balanced (the inverse permutation restores the caller's order)."""
    return np.random.default_rng(seed).permutation(B)


_PERM_CACHE: dict = {}


def _device_permutation(B: int, seed: int | None, device):
    """The shuffle as a tensor on `device`, built once per (B, seed, device): it does not depend on the inputs, and
    generating it on the host and copying it over every step cost more than the result gather itself."""
    import torch

    key = (B, seed, str(device))
    perm = _PERM_CACHE.get(key)
    if perm is None:
        perm = torch.from_numpy(balanced_permutation(B, seed)).to(device) if seed is not None else torch.arange(B, device=device)
        if len(_PERM_CACHE) >= 8:
            _PERM_CACHE.pop(next(iter(_PERM_CACHE)))
        _PERM_CACHE[key] = perm
    return perm


def solve_sharded(solve_local, x0, xs, obs, z_init=None, group=None, shuffle_seed: int | None = 0):
    """Run `solve_local(x0, xs, obs, z_init) -> dict(u0,cost,status,iters)` (torch tensors) on this
    rank's shard and all-gather the results so every rank returns the full batch in input order.
    Works with any torch.distributed backend (nccl on GPUs, gloo in the CPU tests)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    B = x0.shape[0]
    perm = _device_permutation(B, shuffle_seed, x0.device)
    lo, hi = shard_range(B, world, rank)
    idx = perm[lo:hi]
    take = lambda t: None if t is None else t.index_select(0, idx)
    out = solve_local(take(x0), take(xs), take(obs), take(z_init))
    per = -(-B // world)
    n = hi - lo
    if n:
        packed = torch.cat((out["u0"], out["cost"].unsqueeze(1), out["status"].unsqueeze(1).to(torch.float64),
                            out["iters"].unsqueeze(1).to(torch.float64)), dim=1)
        if n < per:  # the padded last shard
            packed = torch.cat((packed, packed.new_zeros((per - n, 5))))
    else:
        packed = torch.zeros((per, 5), dtype=torch.float64, device=x0.device)
    if world > 1:
        full = torch.empty((world * per, 5), dtype=torch.float64, device=x0.device)
        dist.all_gather_into_tensor(full, packed, group=group)
    else:
        full = packed
    # drop the padding of the last shards and undo the shuffle
    if world * per == B:
        rows = full
    else:
        rows = torch.cat([full[r * per: r * per + (shard_range(B, world, r)[1] - shard_range(B, world, r)[0])] for r in range(world)])
    res = torch.empty_like(rows)
    res[perm] = rows
    return {"u0": res[:, 0:2].contiguous(), "cost": res[:, 2].contiguous(), "status": res[:, 3].to(torch.int32),
            "iters": res[:, 4].to(torch.int32)}
