"""world_size-2 (and 3) `gloo` tests of the host-side multi-GPU plumbing (parallelnbody_b200/launch.py); the CUDA/NCCL data
path itself is covered by tests/test_gpu_multi.py on the GPU box."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from parallelnbody_b200 import launch


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert launch.dist_env() == (rank, rank, world)
        uid = launch.broadcast_unique_id(lambda: bytes(range(128)), dist)
        assert uid == bytes(range(128))
        # every rank reports a different "device time"; the bench keeps the max, sums the work
        assert launch.reduce_scalar(10.0 + rank, dist, "max") == 10.0 + world - 1
        assert launch.reduce_scalar(1.5, dist, "sum") == 1.5 * world
        # each rank fills only its slice of a read-back, as nbody_get_positions does
        begin, count, per = launch.partition(n, world, rank)
        truth = np.arange(n * 4, dtype=np.float32).reshape(n, 4)
        mine = np.zeros_like(truth)
        ids = np.arange(begin, begin + count)
        mine[ids] = truth[ids]
        full = launch.combine_shares(mine, ids, n, dist)
        assert np.array_equal(full, truth)
        # overlapping shares are detected
        bad = False
        try:
            launch.combine_shares(truth, np.arange(n), n, dist)
        except RuntimeError:
            bad = True
        assert bad
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 1001), (3, 10)])
def test_host_plumbing_under_gloo(world, n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, "ok") for r in range(world)], res


@pytest.mark.parametrize("n,world", [(0, 1), (1, 1), (10, 3), (1 << 20, 8), (1_000_003, 8), (5, 8)])
def test_partition_tiles_the_bodies(n, world):
    """Same rule as partition() in csrc/nbody_sim.cu: slices are contiguous, disjoint, cover [0, n), equal except the tail."""
    cover = []
    for r in range(world):
        b, c, per = launch.partition(n, world, r)
        assert per == -(-n // world) and 0 <= c <= per
        cover += list(range(b, b + c)) if n < 100 else [(b, c)]
    if n < 100:
        assert cover == list(range(n))
    else:
        assert sum(c for _, c in cover) == n and all(cover[i][0] + cover[i][1] == cover[i + 1][0] for i in range(world - 1))
    with pytest.raises(ValueError):
        launch.partition(10, 2, 2)


def test_unique_id_requires_library_but_not_gpu():
    """ncclGetUniqueId needs no device: rank 0 can mint the id before any CUDA context exists."""
    import parallelnbody_b200 as P
    try:
        a, b = P.comm_unique_id(), P.comm_unique_id()
    except P.NBodyError as e:           # no libnccl on this host: must be the loud NCCL error, never a fallback
        assert e.code == -3
        return
    assert len(a) == 128 and len(b) == 128 and a != b
