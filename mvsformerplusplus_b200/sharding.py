"""Multi-GPU plumbing of the hot path: reference views are independent, so the list of
(scene, ref_view) items is dealt round-robin to one process per GPU; the only collective is the final gather of the
per-view depth / confidence maps (torch.distributed: NCCL on GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_items(n_items, rank, world):
    """Indices of the items rank `rank` processes (round-robin, test.py:238 order preserved within a rank)."""
    return list(range(rank, n_items, world))


def gather_maps(local_maps, n_items, group=None):
    """local_maps [n_local, K, H, W] (K = depth, confidence, ...) for shard_items(n_items, rank, world)
    -> [n_items, K, H, W] in item order on every rank.  Ranks may own different counts (ragged tail)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world == 1:
        return local_maps
    per_rank = (n_items + world - 1) // world
    pad = torch.zeros((per_rank,) + tuple(local_maps.shape[1:]), dtype=local_maps.dtype, device=local_maps.device)
    pad[: local_maps.shape[0]] = local_maps
    buf = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(buf, pad, group=group)
    out = torch.empty((n_items,) + tuple(local_maps.shape[1:]), dtype=local_maps.dtype, device=local_maps.device)
    for r in range(world):
        idx = shard_items(n_items, r, world)
        out[idx] = buf[r][: len(idx)]
    return out
