"""Development probe: where the end-to-end (host buffers in / out) time of a Barnes-Hut step goes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import parallelnbody_b200 as P
from parallelnbody_b200 import ic
from parallelnbody_b200.api import to_particles

n = 1 << 20
posm, vel = ic.plummer(n, seed=1234)
aos_in = torch.empty(n * 40, dtype=torch.uint8).pin_memory()
aos_out = torch.empty(n * 40, dtype=torch.uint8).pin_memory()
aos_in.numpy().view(P.PARTICLE_DTYPE)[:] = to_particles(posm, vel)
for meth in (P.METHOD_BARNES_HUT, P.METHOD_DIRECT):
    with P.OctreeSearch(method=meth, eps=0.01, theta=0.25, PhDeltaTime=1e-3) as s:
        for rep in range(3):
            t0 = time.perf_counter(); s.SetParticlesRaw(aos_in.data_ptr(), n, 40)
            t1 = time.perf_counter(); s.Tick()
            t2 = time.perf_counter(); s.GetParticlesRaw(aos_out.data_ptr(), n, 40)
            t3 = time.perf_counter()
            print(f"method={meth} rep={rep}: set {1e3 * (t1 - t0):.2f} ms  tick {1e3 * (t2 - t1):.2f} ms  get {1e3 * (t3 - t2):.2f} ms  (device step {s.Stats()['ms_last_call']:.2f} ms)", flush=True)
