"""world_size-2 gloo test (CPU) of the N>1 host path: round-robin sharding of reference views and the final gather."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mvsformerplusplus_b200.sharding import gather_maps, shard_items


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_items, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_items(n_items, rank, world)
    local = torch.stack([torch.full((2, 3, 4), float(i)) for i in mine]) if mine else torch.empty(0, 2, 3, 4)
    out = gather_maps(local, n_items)
    ok = all(bool((out[i] == float(i)).all()) for i in range(n_items)) and out.shape[0] == n_items
    q.put((rank, ok))
    dist.destroy_process_group()


def test_shard_items_partition():
    for n in (1, 7, 64):
        for world in (1, 2, 8):
            seen = sorted(i for r in range(world) for i in shard_items(n, r, world))
            assert seen == list(range(n))


def test_gather_maps_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    for n_items in (5, 4):  # ragged (rank 0 owns 3 items, rank 1 owns 2) and even (all_gather_into_tensor + transpose)
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, n_items, q)) for r in range(2)]
        for p in procs:
            p.start()
        res = [q.get(timeout=120) for _ in procs]
        for p in procs:
            p.join(timeout=60)
        assert all(ok for _, ok in res), res
