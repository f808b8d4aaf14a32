"""Multi-GPU plumbing of the hot path: reference views are independent, so the list of
(scene, ref_view) items is dealt round-robin to one process per GPU; the only collective is the final gather of the
per-view depth / confidence maps (torch.distributed: NCCL on GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_items(n_items, rank, world):
    """Indices of the items rank `rank` processes (round-robin, test.py:238 order preserved within a rank)."""
    return list(range(rank, n_items, world))


def gather_maps(local_maps, n_items, group=None, workspace=None):
    """local_maps [n_local, K, H, W] (K = depth, confidence, ...) for shard_items(n_items, rank, world)
    -> [n_items, K, H, W] in item order on every rank.  Ranks may own different counts (ragged tail).
    `workspace` (optional dict, reused across calls) keeps the staging buffers so a steady-state caller does not allocate."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local_maps
    per_rank = (n_items + world - 1) // world
    tail = tuple(local_maps.shape[1:])
    ws = workspace if workspace is not None else {}
    key = (per_rank, tail, local_maps.dtype, str(local_maps.device), world)
    if ws.get("key") != key:
        ws["key"] = key
        ws["all"] = torch.empty((world, per_rank) + tail, dtype=local_maps.dtype, device=local_maps.device)
        ws["pad"] = torch.zeros((per_rank,) + tail, dtype=local_maps.dtype, device=local_maps.device)
        ws["out"] = torch.empty((n_items,) + tail, dtype=local_maps.dtype, device=local_maps.device)
    if local_maps.shape[0] == per_rank:
        src = local_maps.contiguous()
    else:
        src = ws["pad"]
        src[: local_maps.shape[0]] = local_maps
    dist.all_gather_into_tensor(ws["all"].view(-1), src.view(-1), group=group)
    out = ws["out"] if ws["out"].shape[0] == n_items else torch.empty((n_items,) + tail, dtype=local_maps.dtype, device=local_maps.device)
    if n_items == per_rank * world:          # item j * world + r  <-  rank r, slot j
        out.view((per_rank, world) + tail).copy_(ws["all"].transpose(0, 1))
    else:
        for r in range(world):
            idx = shard_items(n_items, r, world)
            out[idx] = ws["all"][r][: len(idx)]
    return out
