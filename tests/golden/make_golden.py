"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/liboracle_ref.so, built by
oracle/Makefile from /root/reference/Source/NBody/OctreeSearch.{h,cpp} against oracle/shim).

Run in the authoring container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py

The reference ships no tests or golden vectors, so these files ARE the pin: the CPU
restatement (oracle/nbody_oracle.c) must reproduce them bit for bit (tests/test_oracle.py), and the CUDA path is
then checked against the restatement / these vectors within the tolerances BASELINE.json states.

Each case stores the injected initial conditions (FParticle AoS, OctreeSearch.h:9-18) and what the reference
computed from them.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402
from parallelnbody_b200 import ic  # noqa: E402

THETAS = [0.0, 0.25, 0.35, 0.5, 1.0]


def forces_case(name, posm, vel):
    """Tree built by CreateOctree (root origin 0, half = ComputeCubeSize), walk at several Theta."""
    r = O.RefSim()
    r.SetParticles(O.to_aos(posm, vel))
    r.ComputeCubeSize()
    r.CreateOctree()                      # OctreeSearch.cpp:74-89 (also walks at the shipped Theta = 1.0)
    out = {"particles0": O.to_aos(posm, vel), "size": np.float32(r.Size)}
    out["acc_shipped"] = r.Particles()["Acceleration"].copy()
    root = r.Root()
    out["root_origin"], out["root_half"] = root["origin"], np.float32(root["half"])
    out["root_mass"], out["root_com"] = np.float32(root["mass"]), root["com"]
    for th in THETAS:
        r.ComputeForces(th)
        out[f"acc_theta_{th:g}"] = r.Particles()["Acceleration"].copy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    r.close()
    print(name, "n =", posm.shape[0], "size =", out["size"])


def tick_case(name, posm, vel, dt, steps, show_octree=True):
    """`steps` calls of AOctreeSearch::Tick as shipped (Theta = 1.0, root origin = previous COM)."""
    r = O.RefSim()
    r.SetParticles(O.to_aos(posm, vel))
    r.PhDeltaTime = dt
    r.set_show_octree(show_octree)
    out = {"particles0": O.to_aos(posm, vel), "dt": np.float32(dt), "steps": np.int32(steps)}
    snaps, roots = [], []
    for _ in range(steps):
        r.Tick(1)
        snaps.append(r.Particles().copy())
        root = r.Root()
        roots.append(np.concatenate([root["origin"], [root["half"], root["mass"]], root["com"]]).astype(np.float32))
    out["particles"] = np.stack(snaps)
    out["roots"] = np.stack(roots)
    out["draws_last"] = r.DebugDraws()    # DrawOctreeBoxes of the last Tick (OctreeSearch.cpp:36-45)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    r.close()
    print(name, "n =", posm.shape[0], "steps =", steps)


def theta0_steps_case(name, posm, vel, dt, steps):
    """The reference's own 'direct sum' trajectory: tree walk at Theta = 0 + its integrator statements."""
    r = O.RefSim()
    r.SetParticles(O.to_aos(posm, vel))
    r.PhDeltaTime = dt
    e = []
    for _ in range(steps):
        r.ComputeCubeSize()
        r.CreateOctree()
        r.ComputeForces(0.0)
        r.Integrate()
    out = {"particles0": O.to_aos(posm, vel), "dt": np.float32(dt), "steps": np.int32(steps), "particles": r.Particles()}
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    r.close()
    print(name, "n =", posm.shape[0], "steps =", steps, e)


if __name__ == "__main__":
    assert O.have_ref(), "needs /root/reference (authoring container)"
    forces_case("slab_256_forces", *ic.reference_slab(256, 1000.0, seed=1234))
    forces_case("plummer_512_forces", *ic.plummer(512, seed=1234))
    forces_case("uniform_1000_forces", *ic.uniform_cube(1000, seed=1234))
    tick_case("slab_300_tick5", *ic.reference_slab(300, 1000.0, seed=7), dt=0.01, steps=5)
    tick_case("plummer_256_tick4", *ic.plummer(256, seed=3), dt=1e-3, steps=4)
    theta0_steps_case("plummer_256_theta0_20steps", *ic.plummer(256, seed=5), dt=1e-3, steps=20)
    # tiny hand-checkable cases
    two = np.array([[0, 0, 0, 2.0], [3, 4, 0, 1.0]], np.float32)
    forces_case("two_body_forces", two, np.zeros_like(two))
    nine = np.array([[x, y, z, 1.0 + 0.5 * i] for i, (x, y, z) in enumerate(
        [(-1, -1, -1), (-1, -1, 1), (-1, 1, -1), (-1, 1, 1), (1, -1, -1), (1, -1, 1), (1, 1, -1), (1, 1, 1), (0.25, 0.5, 0.75)])],
        np.float32)
    forces_case("nine_body_forces", nine, np.zeros_like(nine))
