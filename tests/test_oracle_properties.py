"""Property tests (hypothesis) of the CPU oracle and the host-side partition rule - the size-independent facts the GPU
parity tests lean on: Newton's third law, permutation / translation invariance, linearity in the masses, theta -> 0
limit of the restated tree walk, integrator algebra."""
import numpy as np
from hypothesis import given, settings, strategies as st

from conftest import rel_l2

SET = settings(max_examples=25, deadline=None)


def _bodies(seed, n, spread=1.0):
    rng = np.random.default_rng(seed)
    posm = np.empty((n, 4), np.float32)
    posm[:, :3] = rng.normal(0, spread, (n, 3))
    posm[:, 3] = rng.uniform(0.5, 2.0, n) * 1e-3
    return posm


@SET
@given(seed=st.integers(0, 10_000), n=st.integers(2, 300), eps=st.sampled_from([0.0, 0.05]))
def test_direct_sum_conserves_momentum(oracle, seed, n, eps):
    posm = _bodies(seed, n)
    a = oracle.direct_f64(posm, G=1e4, eps=eps)
    f = (posm[:, 3:4].astype(np.float64) * a).sum(0)
    scale = np.abs(posm[:, 3:4].astype(np.float64) * a).sum(0) + 1e-300
    assert np.all(np.abs(f) <= 1e-10 * scale)


@SET
@given(seed=st.integers(0, 10_000), n=st.integers(2, 200))
def test_direct_sum_permutation_translation_and_mass_linearity(oracle, seed, n):
    posm = _bodies(seed, n)
    a = oracle.direct_f64(posm, eps=0.01)
    perm = np.random.default_rng(seed + 1).permutation(n)
    assert rel_l2(oracle.direct_f64(posm[perm], eps=0.01), a[perm]) <= 1e-13     # fp64 sums in another order
    shifted = posm.copy()
    shifted[:, :3] += np.float32(8.0)          # fp32 coordinates lose 3-4 bits: pair separations change by ~1e-6 relative
    assert rel_l2(oracle.direct_f64(shifted, eps=0.01), a) <= 1e-4
    heavier = posm.copy()
    heavier[:, 3] *= np.float32(4.0)           # exact in fp32
    assert rel_l2(oracle.direct_f64(heavier, eps=0.01), 4.0 * a) <= 1e-15


@SET
@given(seed=st.integers(0, 10_000), n=st.integers(2, 400))
def test_restated_walk_theta0_is_the_direct_sum_and_error_grows_with_theta(oracle, seed, n):
    posm = _bodies(seed, n)
    t = oracle.BHTree(posm, half=oracle.cube_size(posm))
    if t.status != 0:            # coincident fp32 positions: the reference itself would not return
        t.close()
        return
    exact = oracle.direct_f64(posm)
    a0, c0 = t.forces(0.0, return_count=True)
    assert c0 == n * (n - 1) and rel_l2(a0, exact) <= 2e-5
    errs, counts = [], []
    for th in (0.25, 0.5, 1.0):
        a, c = t.forces(th, return_count=True)
        errs.append(rel_l2(a, exact)); counts.append(c)
    assert counts[0] >= counts[1] >= counts[2] and counts[0] <= c0
    t.close()


@SET
@given(seed=st.integers(0, 10_000), n=st.integers(1, 200), dt=st.sampled_from([1e-3, 0.01, 0.5]))
def test_kick_drift_algebra(oracle, seed, n, dt):
    """v' = v + dt*a ; x' = x + dt*v' in fp32 with product-then-add rounding (OctreeSearch.cpp:29-30); mass untouched."""
    rng = np.random.default_rng(seed)
    p = rng.normal(0, 1, (n, 4)).astype(np.float32)
    v = rng.normal(0, 1, (n, 4)).astype(np.float32); v[:, 3] = 0
    a = rng.normal(0, 1, (n, 4)).astype(np.float32); a[:, 3] = 0
    p0, v0 = p.copy(), v.copy()
    oracle.kick_drift(p, v, a, np.float32(dt))
    dtf = np.float32(dt)
    v1 = v0[:, :3] + dtf * a[:, :3]
    x1 = p0[:, :3] + dtf * v1
    assert np.array_equal(v[:, :3], v1) and np.array_equal(p[:, :3], x1) and np.array_equal(p[:, 3], p0[:, 3])


@SET
@given(n=st.integers(0, 5_000_000), world=st.integers(1, 16))
def test_partition_rule(n, world):
    from parallelnbody_b200 import launch
    total, prev_end = 0, 0
    for r in range(world):
        b, c, per = launch.partition(n, world, r)
        assert per == -(-n // world) and b == min(n, r * per) and 0 <= c <= per and b == min(prev_end, n) or c == 0
        prev_end = b + c
        total += c
    assert total == n
