"""Multi-GPU paths. On one GPU: the ranks' slices are emulated by handles that share the device (no collective can run
between processes on one GPU, B200_PROFILING.md); with >= 2 GPUs: the real NCCL path under torchrun (tools/multi_gpu_check.py).
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import rel_l2

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("method", ["direct", "bh"])
def test_emulated_ranks_tile_the_single_gpu_result(world, method):
    import parallelnbody_b200 as P
    from parallelnbody_b200 import ic
    n = 10_007                       # not divisible by the world sizes: the last rank's slice is short
    posm, vel = ic.plummer(n, seed=77)
    meth = P.METHOD_DIRECT if method == "direct" else P.METHOD_BARNES_HUT
    with P.OctreeSearch(method=meth, eps=0.01, theta=0.3) as one:
        one.SetBodies(posm, vel)
        one.CreateOctree()
        want = one.Accelerations()
    got = np.zeros_like(want)
    seen = np.zeros(n, np.int32)
    for r in range(world):
        with P.OctreeSearch(method=meth, eps=0.01, theta=0.3, rank=r, world=world, nccl_unique_id=bytes(128)) as s:
            s.SetBodies(posm, vel)
            s.CreateOctree()
            ids = s.LocalIds()
            st = s.Stats()
            assert st["n_global"] == n and st["n_local"] == len(ids)
            a = s.Accelerations()
            got[ids] = a[ids]
            seen[ids] += 1
    assert np.all(seen == 1)                         # the slices partition the bodies
    if method == "direct":
        assert rel_l2(got, want) <= 2e-6             # same kernel, possibly a different j-split
    else:
        assert np.array_equal(got, want)             # same tree, same groups: bit-identical


def test_real_nccl_ranks_match_single_gpu():
    import torch
    ng = torch.cuda.device_count()
    if ng < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    world = 2 if ng < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(ROOT, "tools", "multi_gpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MULTI-GPU CHECK OK" in r.stdout
