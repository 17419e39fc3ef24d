"""Development probe: Barnes-Hut phase times on one B200 (not a bench line).
usage: bh_timing.py SIZES [sweep]   e.g. bh_timing.py 1048576,16777216   |   bh_timing.py 1048576 sweep"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import parallelnbody_b200 as P
from parallelnbody_b200 import ic

sizes = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1 << 20, 1 << 22, 1 << 24]
sweep = len(sys.argv) > 2 and sys.argv[2] == "sweep"
icname = sys.argv[3] if len(sys.argv) > 3 else "plummer"
for n in sizes:
    posm, vel = ic.make(icname, n, 1234)
    keys = np.random.default_rng(1).integers(0, 1 << 63, n, dtype=np.uint64)
    _, _, ms = P.sort_pairs_u64(keys, 63, timed=True)
    print(f"N={n}: radix sort 63-bit pairs {ms:.3f} ms = {n / ms * 1e-6:.2f} Gkeys/s = {n * 24 * 8 / ms * 1e-6:.0f} GB/s of (12 r + 12 w) x 8 passes")
    combos = [(64, 2, 16)]
    if len(sys.argv) > 2 and sys.argv[2] == "walk":   # the two walk-group shapes that matter, for kernel A/B runs (NBODY_WALK=0/1)
        combos = [(32, 2, 16), (64, 2, 16), (64, 2, 8), (128, 2, 16)]
    if sweep:
        combos = [(gs, pack, leaf) for gs in (32, 64) for pack in (1, 2, 4) for leaf in (8, 16, 32)]
    for th in ((0.25,) if sweep else (0.25, 0.35)):
        for gs, pack, leaf in combos:
            with P.OctreeSearch(method=P.METHOD_BARNES_HUT, eps=0.01, theta=th, leaf_size=leaf, group_size=gs, group_pack=pack) as s:
                s.SetBodies(posm, vel)
                s.Step(1e-3, 3)
                s.Step(1e-3, 10)
                st = s.Stats()
                fill = n / max(1, st['walk_groups']) / gs
                print(f"N={n} theta={th} group={gs} pack={pack} leaf={leaf}: {st['ms_last_call'] / 10:.3f} ms/step  build {st['ms_build'] / 10:.3f}  walk {st['ms_force'] / 10:.3f}  "
                      f"nodes {st['tree_nodes']} depth {st['tree_depth']} passes {st['sort_passes']} groups {st['walk_groups']} fill {fill:.2f}  inter/body {st['interactions'] / n:.0f}  "
                      f"{st['interactions'] / (st['ms_force'] / 10 * 1e-3):.3e} inter/s", flush=True)
