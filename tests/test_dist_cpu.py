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
