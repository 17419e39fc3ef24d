"""Multi-GPU Barnes-Hut (K9: Morton domain split, body migration, locally-essential-tree exchange) on ONE GPU.

The ranks are W handles of this process joined by a loop-back communicator (include/nbody.h, nbody_comm_loopback_id):
same code as under NCCL - splitters, migration all-to-all-v, export descent, LET tree, second walk - with the collectives
done as device-to-device copies. One host thread per rank, as one process per rank under torchrun.

The reference has no multi-GPU counterpart (OctreeSearch.cpp:21-34 is one game-thread loop); the bars are the ones
north_star states for Barnes-Hut: force error against the direct sum no larger than the reference's own error at the
same Theta (1.4e-2 at Theta 0.35, BASELINE.md section 2 - and no larger than the single-GPU walk's), trajectories that follow the
single-GPU run, and exact partition of the bodies over the ranks.
"""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

from conftest import rel_l2

pytestmark = pytest.mark.gpu
os.environ.setdefault("NBODY_LOOPBACK_TIMEOUT_S", "60")
REF_ERR_THETA_035 = 1.4e-2      # the reference's own error at Theta 0.35 (conventional 0.7)


def run_ranks(world, fn):
    """fn(rank, uid) on `world` threads (ctypes releases the GIL inside the library); returns the per-rank results."""
    import parallelnbody_b200 as P
    uid = P.comm_loopback_id()
    with ThreadPoolExecutor(world) as ex:
        futs = [ex.submit(fn, r, uid) for r in range(world)]
        return [f.result() for f in futs]


def combine(n, parts):
    """parts = [(ids, array[n, 4])] with each rank's share filled at its ids -> full array; shares must partition."""
    out = np.zeros((n, 4), np.float32)
    seen = np.zeros(n, np.int32)
    for ids, a in parts:
        out[ids] = a[ids]
        seen[ids] += 1
    assert np.all(seen == 1), "rank shares overlap or miss bodies"
    return out


def _single(posm, vel, theta, eps, dt, steps, method=None):
    import parallelnbody_b200 as P
    with P.OctreeSearch(method=P.METHOD_BARNES_HUT if method is None else method, eps=eps, theta=theta) as s:
        s.SetBodies(posm, vel)
        s.CreateOctree()
        acc = s.Accelerations()
        if steps:
            s.Step(dt, steps)
        return acc, s.Positions(), s.Velocities()


@pytest.mark.parametrize("world", [2, 3, 8])
def test_let_two_galaxies_forces_and_migration(world):
    import parallelnbody_b200 as P
    from parallelnbody_b200 import ic
    n, eps, theta, dt, steps = 60_001, 0.01, 0.35, 0.02, 6
    posm, vel = ic.two_galaxies(n, seed=42)
    with P.OctreeSearch(method=P.METHOD_DIRECT, eps=eps) as d:
        d.SetBodies(posm, vel)
        d.CreateOctree()
        exact = d.Accelerations()
    acc1, pos1, vel1 = _single(posm, vel, theta, eps, dt, steps)

    def rank_fn(r, uid):
        with P.OctreeSearch(method=P.METHOD_BARNES_HUT, eps=eps, theta=theta, rank=r, world=world, nccl_unique_id=uid,
                            bh_exchange=0) as s:
            s.SetBodies(posm, vel)
            ids_set = s.LocalIds()                    # the rows this rank's read-backs fill: its slice of the caller's order
            pos_set = s.Positions()                   # read back before any step: the bodies already live in their domains
            s.CreateOctree()
            ids0 = s.LocalIds()
            acc0 = s.Accelerations()
            st0 = s.Stats()
            moved, sizes = 0, [st0["n_local"]]
            for _ in range(steps):                    # one step per call so that every step's migration count is seen
                s.Step(dt, 1)
                st = s.Stats()
                moved += st["migrated"]
                sizes.append(st["n_local"])
            ids = s.LocalIds()
            return dict(ids_set=ids_set, pos_set=pos_set, ids0=ids0, acc0=acc0, st0=st0, ids=ids, pos=s.Positions(),
                        vel=s.Velocities(), acc=s.Accelerations(), st=s.Stats(), moved=moved, sizes=sizes)

    out = run_ranks(world, rank_fn)
    # read-backs right after the upload land at the bodies' own rows (each rank holds a slice of the caller's order)
    got_set = combine(n, [(o["ids_set"], o["pos_set"]) for o in out])
    assert np.array_equal(got_set, posm)
    # forces: the LET walk accepts remote cells more strictly than the single-GPU walk, never more loosely
    acc = combine(n, [(o["ids0"], o["acc0"]) for o in out])
    err_let, err_one = rel_l2(acc, exact), rel_l2(acc1, exact)
    assert err_let <= REF_ERR_THETA_035
    assert err_let <= err_one * 1.05 + 1e-6, f"LET {err_let:.3e} vs single GPU {err_one:.3e}"
    assert all(o["st0"]["let_points"] > 0 for o in out)
    assert sum(o["st0"]["n_local"] for o in out) == n
    # trajectories after steps in which bodies changed owner
    pos = combine(n, [(o["ids"], o["pos"]) for o in out])
    velc = combine(n, [(o["ids"], o["vel"]) for o in out])
    assert rel_l2(pos, pos1) <= 1e-4 and rel_l2(velc, vel1) <= 5e-3   # two Theta-0.35 walks, 6 steps of dt 0.02
    assert np.array_equal(pos[:, 3], posm[:, 3])                       # masses travelled with their bodies
    assert all(np.array_equal(o["ids"], o["ids_set"]) for o in out)   # read-backs return every body to its owner's rows
    moved = sum(o["moved"] for o in out)
    assert moved > 0, "no body migrated: the test does not exercise the migration path"
    for k in range(steps + 1):
        assert sum(o["sizes"][k] for o in out) == n                   # no body lost or duplicated in any step
    acc_end = combine(n, [(o["ids"], o["acc"]) for o in out])
    assert np.all(np.isfinite(acc_end))


def test_let_theta0_is_the_direct_sum():
    """Theta = 0: nothing is ever accepted, every rank imports every other body - the exchange must deliver exactly them."""
    import parallelnbody_b200 as P
    from parallelnbody_b200 import ic
    n, eps, world = 6000, 0.02, 3
    posm, vel = ic.two_galaxies(n, seed=7)
    with P.OctreeSearch(method=P.METHOD_DIRECT, eps=eps) as d:
        d.SetBodies(posm, vel)
        d.CreateOctree()
        exact = d.Accelerations()

    def rank_fn(r, uid):
        with P.OctreeSearch(method=P.METHOD_BARNES_HUT, eps=eps, theta=0.0, rank=r, world=world, nccl_unique_id=uid,
                            bh_exchange=0) as s:
            s.SetBodies(posm, vel)
            s.CreateOctree()
            st = s.Stats()
            return s.LocalIds(), s.Accelerations(), st

    out = run_ranks(world, rank_fn)
    acc = combine(n, [(o[0], o[1]) for o in out])
    assert rel_l2(acc, exact) <= 1e-5
    for _, _, st in out:
        assert st["let_points"] == n - st["n_local"]
    assert sum(o[2]["interactions"] for o in out) == float(n) * n


def test_let_energy_and_particles_roundtrip():
    """Energy is a collective over the ranks; FParticle read-back (Particles) fills each rank's share."""
    import parallelnbody_b200 as P
    from parallelnbody_b200 import ic
    n, eps, world = 20_000, 0.01, 4
    posm, vel = ic.plummer(n, seed=3)
    with P.OctreeSearch(method=P.METHOD_BARNES_HUT, eps=eps, theta=0.3) as one:
        one.SetBodies(posm, vel)
        e_one = one.Energy()
        one.Step(1e-3, 4)
        part_one = one.Particles

    def rank_fn(r, uid):
        with P.OctreeSearch(method=P.METHOD_BARNES_HUT, eps=eps, theta=0.3, rank=r, world=world, nccl_unique_id=uid,
                            bh_exchange=0) as s:
            s.Particles = P.api.to_particles(posm, vel)
            e = s.Energy()
            s.Step(1e-3, 4)
            return e, s.LocalIds(), s.Particles

    out = run_ranks(world, rank_fn)
    for e, _, _ in out:
        assert abs(e[0] - e_one[0]) <= 1e-9 * abs(e_one[0]) and abs(e[1] - e_one[1]) <= 1e-6 * abs(e_one[1])
    full = np.zeros(n, P.PARTICLE_DTYPE)
    seen = np.zeros(n, np.int32)
    for _, ids, part in out:
        full[ids] = part[ids]
        seen[ids] += 1
    assert np.all(seen == 1)
    assert np.array_equal(full["Mass"], part_one["Mass"])
    assert rel_l2(full["Position"], part_one["Position"]) <= 1e-6
    assert rel_l2(full["Velocity"], part_one["Velocity"]) <= 1e-4


@pytest.mark.parametrize("method,exchange", [("direct", 0), ("bh", 1)])
def test_loopback_direct_and_replicated_are_bit_identical(method, exchange):
    """The all-gather paths (direct sum i-rows; replicated-tree Barnes-Hut) under the loop-back communicator."""
    import parallelnbody_b200 as P
    from parallelnbody_b200 import ic
    n, eps, world = 30_011, 0.01, 3
    posm, vel = ic.plummer(n, seed=19)
    meth = P.METHOD_DIRECT if method == "direct" else P.METHOD_BARNES_HUT
    _, pos1, vel1 = _single(posm, vel, 0.3, eps, 1e-3, 5, method=meth)

    def rank_fn(r, uid):
        with P.OctreeSearch(method=meth, eps=eps, theta=0.3, rank=r, world=world, nccl_unique_id=uid, bh_exchange=exchange) as s:
            s.SetBodies(posm, vel)
            s.CreateOctree()          # as _single does: the second build then sorts the same number of key levels
            s.Step(1e-3, 5)
            return s.LocalIds(), s.Positions(), s.Velocities()

    out = run_ranks(world, rank_fn)
    pos = combine(n, [(o[0], o[1]) for o in out])
    velc = combine(n, [(o[0], o[2]) for o in out])
    if method == "bh":
        assert np.array_equal(pos, pos1) and np.array_equal(velc, vel1)      # same tree, same groups
    else:
        assert rel_l2(pos, pos1) <= 1e-7 and rel_l2(velc, vel1) <= 2e-6      # same kernel, another j-split


def test_let_upload_tick_readback_every_frame():
    """The end-to-end pattern of a host application (and of bench.py's e2e leg): FParticle array in, Tick, FParticle array
    out, every frame. Each upload re-uses the domains of the frame before (kept splitters and root cube); results follow
    a single GPU doing the same."""
    import parallelnbody_b200 as P
    from parallelnbody_b200 import ic
    n, eps, world, frames = 40_003, 0.01, 4, 4
    posm, vel = ic.two_galaxies(n, seed=11)
    start = P.api.to_particles(posm, vel)
    with P.OctreeSearch(method=P.METHOD_BARNES_HUT, eps=eps, theta=0.3, PhDeltaTime=0.01) as one:
        cur = start.copy()
        for _ in range(frames):
            one.Particles = cur
            one.Tick()
            cur = one.Particles
        want = cur

    def rank_fn(r, uid):
        with P.OctreeSearch(method=P.METHOD_BARNES_HUT, eps=eps, theta=0.3, PhDeltaTime=0.01, rank=r, world=world,
                            nccl_unique_id=uid, bh_exchange=0) as s:
            cur = start.copy()
            for _ in range(frames):
                s.Particles = cur
                s.Tick()
                mine = s.Particles
                ids = s.LocalIds()
                # a real multi-process caller would all-gather the shares; the threads of this test share the array
                shared[ids] = mine[ids]
                barrier.wait()
                cur = shared.copy()
                barrier.wait()
            return ids

    import threading
    shared = np.zeros(n, P.PARTICLE_DTYPE)
    barrier = threading.Barrier(world)
    out = run_ranks(world, rank_fn)
    assert sum(len(i) for i in out) == n
    assert np.array_equal(shared["Mass"], want["Mass"])
    assert rel_l2(shared["Position"], want["Position"]) <= 1e-5
    assert rel_l2(shared["Velocity"], want["Velocity"]) <= 5e-3
    assert rel_l2(shared["Acceleration"], want["Acceleration"]) <= 2e-2     # two Theta-0.3 evaluations of the same field
