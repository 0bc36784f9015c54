"""Multi-GPU sharding of the localisation hot path (one process per GPU).

Tomograms are independent (the reference loops them serially, cet_pick/test.py:82-85), so the list is
split by tomogram over the ranks and there is NO data-path collective; the only exchange is the
gather of the (n_local, K, 5) pick tensors.  Works over NCCL (CUDA tensors) and gloo (CPU tensors)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous share [first, first + count) of `n_items` for `rank`; the first n % world ranks get one more."""
    if world <= 0 or not (0 <= rank < world) or n_items < 0:
        raise ValueError("bad shard arguments")
    base, extra = divmod(n_items, world)
    count = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return first, count


def gather_picks(local: torch.Tensor, n_items: int, group=None) -> torch.Tensor:
    """All ranks call this with their (count_r, K, 5) picks (shard_range order); every rank gets the
    (n_items, K, 5) tensor in tomogram order.  Unequal shares are padded to the largest share for the
    single all_gather and trimmed afterwards."""
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    K, C = local.shape[1], local.shape[2]
    max_count = -(-n_items // world)
    buf = local.new_zeros((max_count, K, C))
    buf[:local.shape[0]] = local
    out = local.new_empty((world * max_count, K, C))
    dist.all_gather_into_tensor(out, buf.contiguous(), group=group)
    parts = []
    for r in range(world):
        _, cnt = shard_range(n_items, r, world)
        parts.append(out[r * max_count:r * max_count + cnt])
    return torch.cat(parts, 0)
