"""Development probe (1 GPU): the domain-split Barnes-Hut path with W loop-back ranks - per-rank body counts, imported
locally-essential points, migration, after a few steps. Timings share one GPU and are only indicative.
usage: let_probe.py N W [steps] [ic]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from concurrent.futures import ThreadPoolExecutor
import numpy as np
import parallelnbody_b200 as P
from parallelnbody_b200 import ic

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
icname = sys.argv[4] if len(sys.argv) > 4 else "two_galaxies"
posm, vel = ic.make(icname, n, 1234)
uid = P.comm_loopback_id()

def rank_fn(r):
    rows = []
    with P.OctreeSearch(method=P.METHOD_BARNES_HUT, eps=0.01, theta=0.35, rank=r, world=world, nccl_unique_id=uid, bh_exchange=0) as s:
        s.SetBodies(posm, vel)
        for k in range(steps):
            s.Step(1e-3, 1)
            st = s.Stats()
            rows.append((st["n_local"], st["let_points"], st["migrated"], st["interactions"], st["ms_last_call"], st["ms_build"], st["ms_force"]))
    return rows

t0 = time.time()
with ThreadPoolExecutor(world) as ex:
    out = list(ex.map(rank_fn, range(world)))
print(f"N={n} W={world} {icname}: {time.time() - t0:.1f} s")
for k in range(steps):
    nl = [o[k][0] for o in out]; let = [o[k][1] for o in out]; mig = [o[k][2] for o in out]; it = [o[k][3] for o in out]
    print(f"step {k}: n_local {min(nl)}..{max(nl)}  let_points {min(let)}..{max(let)} (sum {sum(let)})  migrated {sum(mig)}  "
          f"interactions/rank {min(it):.3e}..{max(it):.3e} (imbalance {max(it) / (sum(it) / world):.2f})  ms/step(shared GPU) {max(o[k][4] for o in out):.2f}")
