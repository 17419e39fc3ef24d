"""Small end-to-end run for `compute-sanitizer --tool memcheck` (one tool per gpurun call, tiny sizes)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import parallelnbody_b200 as P
from parallelnbody_b200 import ic

posm, vel = ic.plummer(3001, seed=1)
with P.OctreeSearch(method=P.METHOD_DIRECT, eps=0.01) as s:
    s.SetBodies(posm, vel); s.Step(1e-3, 2); s.Energy(); p = s.Particles
with P.OctreeSearch(method=P.METHOD_DIRECT, eps=0.0) as s:
    s.CreateSpacePoints(777, 100.0); s.Tick()
posm, vel = ic.plummer(20011, seed=2)
posm[100:400, :3] = posm[7, :3]
for kw in (dict(), dict(mac=1, leaf_size=1, reference_root=True), dict(group_size=32, group_pack=4, leaf_size=4), dict(group_size=128)):
    with P.OctreeSearch(method=P.METHOD_BARNES_HUT, eps=0.01, theta=0.4, **kw) as s:
        s.SetBodies(posm, vel); s.Step(1e-3, 3); s.OctreeBoxes(); s.OctreeNodes(); s.Energy(); s.Positions(); p = s.Particles
with P.OctreeSearch(method=P.METHOD_BARNES_HUT, theta=0.5) as s:
    s.SetBodies(posm[:1]); s.Tick(); s.SetBodies(posm[:70]); s.Tick()
k, i = P.sort_pairs_u64(np.random.default_rng(0).integers(0, 1 << 63, 10007, dtype=np.uint64), 63)
# domain-split Barnes-Hut (migration + LET exchange + second walk) over the loop-back communicator: 3 ranks, one thread each
from concurrent.futures import ThreadPoolExecutor
posm, vel = ic.two_galaxies(9001, seed=3)
uid = P.comm_loopback_id()
def rank_fn(r):
    with P.OctreeSearch(method=P.METHOD_BARNES_HUT, eps=0.01, theta=0.35, rank=r, world=3, nccl_unique_id=uid, bh_exchange=0) as s:
        s.SetBodies(posm, vel); s.Step(0.02, 3); s.Energy(); s.Positions(); return s.Stats()["let_points"]
with ThreadPoolExecutor(3) as ex:
    print("let points", list(ex.map(rank_fn, range(3))))
print("sanitize probe done")
