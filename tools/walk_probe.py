"""Development probe for ncu: a few Barnes-Hut steps at one walk-group shape.  usage: walk_probe.py N GROUP [THETA] [STEPS]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import parallelnbody_b200 as P
from parallelnbody_b200 import ic

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
gs = int(sys.argv[2]) if len(sys.argv) > 2 else 64
th = float(sys.argv[3]) if len(sys.argv) > 3 else 0.25
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 4
posm, vel = ic.make("plummer", n, 1234)
with P.OctreeSearch(method=P.METHOD_BARNES_HUT, eps=0.01, theta=th, leaf_size=16, group_size=gs, group_pack=2) as s:
    s.SetBodies(posm, vel)
    s.Step(1e-3, 2)
    s.Step(1e-3, steps)
    st = s.Stats()
    print(f"N={n} group={gs} theta={th}: walk {st['ms_force'] / steps:.3f} ms  build {st['ms_build'] / steps:.3f} ms  inter/body {st['interactions'] / n:.0f}")
