"""Host-side plumbing for the multi-GPU paths: one process per GPU under torchrun, ``torch.distributed`` only for the
out-of-band pieces (broadcasting the NCCL id the C++ library bootstraps its own communicator from, reducing timings,
combining read-backs). The data path - the per-step all-gather of float4 positions - runs inside libnbody_b200.so on
NCCL directly (csrc/comm.cpp), not through torch.

``partition`` mirrors ``partition()`` in csrc/nbody_sim.cu: contiguous slices of ceil(n / world) bodies; for the direct
sum the slice is a range of original indices, for Barnes-Hut a range of the Morton order.
"""
from __future__ import annotations

import os

import numpy as np


def dist_env():
    """(rank, local_rank, world) from the torchrun environment (1 process = 1 GPU)."""
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def partition(n: int, world: int, rank: int):
    """(begin, count, per) of rank's slice: per = ceil(n / world), begin = min(n, rank * per), count = min(per, n - begin)."""
    if world < 1 or not 0 <= rank < world or n < 0:
        raise ValueError("bad partition arguments")
    per = -(-n // world)
    begin = min(n, rank * per)
    return begin, min(per, n - begin), per


def broadcast_unique_id(make_id, dist, device=None) -> bytes:
    """Rank 0 calls make_id() (-> 128 bytes, ``comm_unique_id``); every rank returns the same bytes."""
    import torch
    rank = dist.get_rank()
    t = torch.zeros(128, dtype=torch.uint8, device=device)
    if rank == 0:
        b = make_id()
        if len(b) != 128:
            raise ValueError("an ncclUniqueId is 128 bytes")
        t.copy_(torch.frombuffer(bytearray(b), dtype=torch.uint8))
    dist.broadcast(t, 0)
    return bytes(t.cpu().numpy().tobytes())


def reduce_scalar(x: float, dist, op: str = "max", device=None) -> float:
    """max / sum of a host scalar over the ranks (timings are reported as the max over ranks)."""
    import torch
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(x)
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
    return float(t.item())


def combine_shares(local: np.ndarray, ids: np.ndarray, n: int, dist, device=None) -> np.ndarray:
    """Each rank holds valid rows ``local[ids]`` of an [n, k] array (the get_* calls fill only the rank's share);
    returns the full array on every rank and checks that the shares are disjoint and complete."""
    import torch
    full = torch.zeros((n, local.shape[1]), dtype=torch.float32, device=device)
    cnt = torch.zeros(n, dtype=torch.int32, device=device)
    idx = torch.from_numpy(np.asarray(ids, np.int64)).to(full.device)
    full[idx] = torch.from_numpy(np.ascontiguousarray(local[ids], np.float32)).to(full.device)
    cnt[idx] += 1
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(full)
        dist.all_reduce(cnt)
    if not bool((cnt == 1).all().item()):
        raise RuntimeError("rank shares overlap or miss bodies")
    return full.cpu().numpy()
