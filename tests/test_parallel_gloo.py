"""CPU test of the N>1 plumbing with world_size=2 over gloo: shard ranges, the descriptor-block all-gather
(BASELINE config 4's only exchange step) and the bench's max-time / sum-units reduction."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from orb_slam_b200 import KP_DTYPE, parallel as P


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nf = 50
    rng = np.random.default_rng(100 + rank)
    count = 30 + 7 * rank
    kps = np.zeros(nf, KP_DTYPE)
    kps["x"][:count] = rng.uniform(0, 640, count)
    kps["octave"][:count] = rank
    desc = rng.integers(0, 256, (nf, 32), dtype=np.uint8)
    blk = P.pack_block(torch, torch.from_numpy(kps.view(np.uint8).reshape(nf, 28)), torch.from_numpy(desc), count, nf, "cpu")
    allb = P.allgather_blocks(torch, dist, blk).numpy()
    ok = allb.shape == (world, P.block_bytes(nf))
    for r in range(world):
        k, d = P.unpack_block(allb[r], KP_DTYPE)
        ok &= len(k) == 30 + 7 * r and bool(np.all(k["octave"] == r)) and d.shape == (30 + 7 * r, 32)
        if r == rank:
            ok &= np.array_equal(d, desc[:count]) and np.array_equal(k, kps[:count])
    ms, (kp, nm) = P.reduce_timing(torch, dist, 10.0 + rank, [1000 * (rank + 1), 5], "cpu")
    ok &= ms == 10.0 + world - 1 and kp == 1000 * world * (world + 1) / 2 and nm == 5 * world
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_gloo_world2():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        assert dict(ret) == {0: True, 1: True}


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 10000):
        for world in (1, 2, 4, 8):
            spans = [P.shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
