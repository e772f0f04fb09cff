"""CPU test of the N>1 host logic: pair sharding and the transform all-gather, world_size 2 over gloo."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gaussreg_b200 import parallel


def _fake_transform(pair_index):
    T = torch.eye(4)
    T[:3, 3] = torch.tensor([pair_index, 2.0 * pair_index, -1.0 * pair_index])
    T[0, 1] = 0.001 * pair_index
    return T


def _worker(rank, world, port, n_pairs, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = parallel.shard_pairs(n_pairs, rank, world)
    local = torch.stack([_fake_transform(i) for i in mine]) if mine else torch.zeros((0, 4, 4))
    full = parallel.gather_transforms(local, n_pairs)
    ret[rank] = full
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_and_gather_world2():
    for n_pairs in (7, 8, 1):
        mgr = mp.Manager()
        ret = mgr.dict()
        port = _free_port()
        mp.spawn(_worker, args=(2, port, n_pairs, ret), nprocs=2, join=True)
        want = torch.stack([_fake_transform(i) for i in range(n_pairs)])
        for r in range(2):
            assert torch.equal(ret[r], want), (n_pairs, r)


def test_shard_pairs_partition():
    for world in (1, 2, 4, 8):
        seen = sorted(i for r in range(world) for i in parallel.shard_pairs(1024, r, world))
        assert seen == list(range(1024))
        assert all(parallel.local_count(1024, r, world) == 1024 // world for r in range(world))
    assert parallel.gather_transforms(torch.eye(4)[None], 1, rank=0, world=1).shape == (1, 4, 4)
