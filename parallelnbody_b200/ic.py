"""Synthetic initial conditions (harness, numpy). Seeded with a counter-based generator (Philox) so every
rank / the CPU oracle / the CUDA path see bit-identical bodies.

Units follow SURVEY.md §8d: the reference bakes G = 1e4 (OctreeSearch.h:104), so total mass M = 1e-4 gives
G*M = 1 (standard N-body units) and equal masses m = M/N.

Layout returned: ``posm`` float32[N,4] = (x, y, z, mass) and ``vel`` float32[N,4] = (vx, vy, vz, 0) - the SoA
float4 layout the CUDA path keeps in HBM. ``reference_slab`` reproduces the shape of the reference's own
generator, AOctreeSearch::CreateSpacePoints (OctreeSearch.cpp:58-72).
"""
from __future__ import annotations

import numpy as np

G_REF = 1.0e4
M_TOTAL = 1.0e-4


def _rng(seed: int, stream: int = 0) -> np.random.Generator:
    return np.random.Generator(np.random.Philox(key=[seed, stream]))


def _iso(rng, n):
    """n isotropic unit vectors (float64)."""
    z = rng.uniform(-1.0, 1.0, n)
    phi = rng.uniform(0.0, 2.0 * np.pi, n)
    s = np.sqrt(np.maximum(0.0, 1.0 - z * z))
    return np.stack([s * np.cos(phi), s * np.sin(phi), z], axis=1)


def _pack(pos, vel, mass):
    n = pos.shape[0]
    posm = np.empty((n, 4), np.float32)
    v4 = np.zeros((n, 4), np.float32)
    posm[:, :3] = pos
    posm[:, 3] = mass
    v4[:, :3] = vel
    return posm, v4


def uniform_cube(n: int, seed: int = 1234, half: float = 1.0, total_mass: float = M_TOTAL):
    """BASELINE config 2: uniform random cube [-half, half)^3, zero velocities, equal masses."""
    rng = _rng(seed, 1)
    pos = rng.uniform(-half, half, (n, 3))
    return _pack(pos, np.zeros((n, 3)), np.full(n, total_mass / n))


def plummer(n: int, seed: int = 1234, a: float = 1.0, rmax: float = 10.0, total_mass: float = M_TOTAL,
            G: float = G_REF, stream: int = 2):
    """Plummer sphere in virial equilibrium (Aarseth, Henon & Wielen 1974 sampling), truncated at r < rmax*a,
    centred on its centre of mass with zero net momentum. BASELINE configs 1, 3, 4."""
    rng = _rng(seed, stream)
    r = np.empty(n)
    filled = 0
    while filled < n:
        x = rng.uniform(0.0, 1.0, n - filled)
        x = x[x > 1e-12]
        rr = 1.0 / np.sqrt(np.maximum(x ** (-2.0 / 3.0) - 1.0, 1e-300))
        rr = rr[rr < rmax]
        r[filled:filled + rr.size] = rr
        filled += rr.size
    pos = _iso(rng, n) * (a * r)[:, None]
    q = np.empty(n)
    filled = 0
    while filled < n:
        m = n - filled
        qq = rng.uniform(0.0, 1.0, 2 * m)
        yy = rng.uniform(0.0, 0.1, 2 * m)
        ok = qq[yy < qq * qq * (1.0 - qq * qq) ** 3.5][:m]
        q[filled:filled + ok.size] = ok
        filled += ok.size
    gm = G * total_mass
    vesc = np.sqrt(2.0 * gm / a) * (1.0 + r * r) ** (-0.25)
    vel = _iso(rng, n) * (q * vesc)[:, None]
    pos -= pos.mean(axis=0)
    vel -= vel.mean(axis=0)
    return _pack(pos, vel, np.full(n, total_mass / n))


def two_galaxies(n: int, seed: int = 1234, sep: float = 4.0, vx: float = 0.5, vy: float = 0.2,
                 total_mass: float = M_TOTAL):
    """BASELINE config 5: two Plummer spheres of n/2 bodies, centres (+-sep, 0, 0), approach velocities
    (-+vx, +-vy, 0)."""
    n1 = n // 2
    n2 = n - n1
    p1, v1 = plummer(n1, seed, total_mass=total_mass * n1 / n, stream=3)
    p2, v2 = plummer(n2, seed, total_mass=total_mass * n2 / n, stream=4)
    p1[:, 0] += sep
    p2[:, 0] -= sep
    v1[:, 0] -= vx
    v1[:, 1] += vy
    v2[:, 0] += vx
    v2[:, 1] -= vy
    return np.concatenate([p1, p2]), np.concatenate([v1, v2])


def reference_slab(n: int, size: float = 1000.0, seed: int = 1234):
    """Same distribution as AOctreeSearch::CreateSpacePoints (OctreeSearch.cpp:58-72): uniform slab
    +-(S, S, S/10), speed 10*randint(25..50) in a random direction, mass randint(1..5000); body 0 is a
    5000-mass body at rest at the origin. (The reference draws from the unseeded C rand(); here the draw
    is seeded, so bodies differ from any particular reference run but follow the same law.)"""
    rng = _rng(seed, 5)
    pos = rng.uniform(-1.0, 1.0, (n, 3)) * np.array([size, size, size / 10.0])
    speed = 10.0 * rng.integers(25, 51, n)
    vel = _iso(rng, n) * speed[:, None]
    mass = rng.integers(1, 5001, n).astype(np.float64)
    pos[0] = 0.0
    vel[0] = 0.0
    mass[0] = 5000.0
    return _pack(pos, vel, mass)


def make(name: str, n: int, seed: int = 1234):
    if name == "plummer":
        return plummer(n, seed)
    if name == "uniform":
        return uniform_cube(n, seed)
    if name == "two_galaxies":
        return two_galaxies(n, seed)
    if name == "slab":
        return reference_slab(n, seed=seed)
    raise ValueError(f"unknown initial condition {name!r}")
