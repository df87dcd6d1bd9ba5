"""world_size-2 gloo test of the frame sharding + final gather (host logic of the multi-GPU path)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bodyfitting_b200.sharding import frame_range, gather_frames


def test_frame_range_is_a_balanced_partition():
    for n, w in ((10000, 8), (10, 4), (7, 8), (1, 2), (0, 3)):
        rs = [frame_range(n, r, w) for r in range(w)]
        assert rs[0][0] == 0 and rs[-1][1] == n
        assert all(rs[i][1] == rs[i + 1][0] for i in range(w - 1))
        sizes = [hi - lo for lo, hi in rs]
        assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    lo, hi = frame_range(n, rank, world)
    local = torch.arange(lo, hi, dtype=torch.float32)[:, None] * torch.ones(1, 5)
    full = gather_frames(local, n)
    q.put((rank, full.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_frames_world2_ragged():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    n, world = 7, 2
    ps = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    [p.start() for p in ps]
    got = [q.get(timeout=120) for _ in range(world)]
    [p.join(timeout=60) for p in ps]
    want = np.arange(n, dtype=np.float32)[:, None] * np.ones((1, 5), np.float32)
    for rank, arr in got:
        assert np.array_equal(arr, want), rank


def _halo_worker(rank, world, port, q):
    from bodyfitting_b200.sharding import exchange_halo
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    first = torch.full((5,), 10.0 * rank)          # this shard's first / last frame rows
    last = torch.full((5,), 10.0 * rank + 1)
    prev, nxt = exchange_halo(first, last)
    q.put((rank, None if prev is None else prev.numpy(), None if nxt is None else nxt.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_halo_exchange_world3():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    world = 3
    ps = [ctx.Process(target=_halo_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in ps]
    got = {r: (a, b) for r, a, b in [q.get(timeout=120) for _ in range(world)]}
    [p.join(timeout=60) for p in ps]
    assert got[0][0] is None and got[world - 1][1] is None
    for r in range(world):
        if r > 0:
            assert np.array_equal(got[r][0], np.full(5, 10.0 * (r - 1) + 1, np.float32))   # last row of rank r-1
        if r + 1 < world:
            assert np.array_equal(got[r][1], np.full(5, 10.0 * (r + 1), np.float32))       # first row of rank r+1
